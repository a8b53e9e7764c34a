/* multi_driver.c — drives liblwfb90's multi-GPU entry points from plain C (no Python, no torch, no NCCL): what a Julia
 * host would do with ccall.  Test infrastructure.
 *
 *   multi_driver <blocks.bin> <n_shards> [balanced]
 *
 * blocks.bin (written by tests/c/make_blocks.py): int32 N, iso, n_days, n_rows, M; forcing [n_rows][16]; scal [M][PS];
 * lay [M][PL][N]; u0 [M][L]; expected final state [M][L] of a single-handle run.  Exit code 0 iff the sharded run on
 * devices 0 .. n_shards-1 reproduces the expected final states bit for bit. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lwfb90.h"

static void* rd(FILE* f, size_t n) {
  void* p = malloc(n);
  if (!p || fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s blocks.bin n_shards [balanced]\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t hd[5];
  if (fread(hd, sizeof(int32_t), 5, f) != 5) return 2;
  const int N = hd[0], iso = hd[1], n_days = hd[2], n_rows = hd[3];
  const int64_t M = hd[4];
  const int L = LWFB_STATE_LEN(N, iso);
  double* forcing = rd(f, (size_t)n_rows * LWFB_F * 8);
  double* scal = rd(f, (size_t)M * LWFB_PS * 8);
  double* lay = rd(f, (size_t)M * LWFB_PL * N * 8);
  double* u0 = rd(f, (size_t)M * L * 8);
  double* want = rd(f, (size_t)M * L * 8);
  fclose(f);
  const int n_shards = atoi(argv[2]);
  const int balanced = argc > 3;
  int32_t dev[64];
  for (int s = 0; s < n_shards; s++) dev[s] = s;
  lwfb_opts o;
  lwfb_default_opts(&o);
  o.n_layers = N; o.simulate_isotopes = iso; o.n_days = n_days; o.record_daily = 0;
  lwfb_multi* mh = NULL;
  int rc = lwfb_multi_create(&mh, &o, n_shards, dev);
  if (rc) { fprintf(stderr, "lwfb_multi_create: %d\n", rc); return 1; }
#define OK(call) do { rc = (call); if (rc) { fprintf(stderr, "%s: %d %s\n", #call, rc, lwfb_multi_last_error(mh)); return 1; } } while (0)
  OK(lwfb_multi_set_sites(mh, 1, n_rows, forcing));
  int32_t* site = calloc((size_t)M, sizeof(int32_t));
  double* cost = NULL;
  if (balanced) {          /* cost estimate: the RHS evaluations of a short probe would go here; any numbers do for the test */
    cost = malloc((size_t)M * 8);
    for (int64_t m = 0; m < M; m++) cost[m] = (double)((m * 2654435761u) % 1000);
  }
  OK(lwfb_multi_set_members(mh, M, site, scal, lay, u0, cost));
  OK(lwfb_multi_run(mh));
  double* got = malloc((size_t)M * L * 8);
  OK(lwfb_multi_get_final(mh, got));
  float ms[64], mx;
  OK(lwfb_multi_last_run_ms(mh, ms, &mx));
  long bad = 0;
  for (size_t k = 0; k < (size_t)M * L; k++)
    if (memcmp(got + k, want + k, 8) != 0 && !(got[k] != got[k] && want[k] != want[k])) bad++;
  printf("%lld members x %d days on %d shard(s)%s: slowest shard %.1f ms, %ld state entries differ from the single-handle run\n",
         (long long)M, n_days, n_shards, balanced ? " (balanced)" : "", mx, bad);
  lwfb_multi_destroy(mh);
  return bad != 0;
}
