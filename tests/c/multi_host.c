/* multi_host.c — a C host of liblwfb90 (no Python, no torch, no NCCL in the process): what a Julia / C / Fortran caller does.
 *
 * TEST PROGRAM (tests/test_gpu_parity.py::test_c_host_drives_the_multi_gpu_abi builds it with gcc and runs it).
 * Reads the blocks of an ensemble from a binary file, shards them over the given devices with lwfb_multi_*, integrates,
 * all-gathers the scores peer to peer and writes final states, status, counters and scores to a second file.
 *
 * file layout (little endian): int32 N, iso, t0_day, n_days, n_rows, n_obs_days, n_obs_layers, pad; int64 M;
 *   forcing [n_rows][LWFB_F], site [M] int32 (+ pad to 8), scal [M][LWFB_PS], lay [M][LWFB_PL][N], u0 [M][L], cost [M],
 *   obs_day [n_obs_days] int32, obs_layer [n_obs_layers] int32 (+ pad to 8), obs_theta, obs_d18o [n_obs_days][n_obs_layers]
 * usage: multi_host in.bin out.bin device [device ...]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/lwfb90.h"

static void* rd(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 8);
  if (!p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
  if (bytes % 8) fseek(f, (long)(8 - bytes % 8), SEEK_CUR);
  return p;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    int rc_ = (call);                                                                         \
    if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mh ? lwfb_multi_last_error(mh) : ""); return 3; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: multi_host in.bin out.bin device [device ...]\n"); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t hd[8];
  int64_t M;
  if (fread(hd, 4, 8, f) != 8 || fread(&M, 8, 1, f) != 1) return 2;
  const int N = hd[0], iso = hd[1], t0 = hd[2], nd = hd[3], n_rows = hd[4], n_od = hd[5], n_ol = hd[6];
  const int L = LWFB_STATE_LEN(N, iso);
  double* forcing = rd(f, (size_t)n_rows * LWFB_F * 8);
  int32_t* site = rd(f, (size_t)M * 4);
  double* scal = rd(f, (size_t)M * LWFB_PS * 8);
  double* lay = rd(f, (size_t)M * LWFB_PL * N * 8);
  double* u0 = rd(f, (size_t)M * L * 8);
  double* cost = rd(f, (size_t)M * 8);
  int32_t* obs_day = rd(f, (size_t)n_od * 4);
  int32_t* obs_layer = rd(f, (size_t)n_ol * 4);
  double* obs_theta = rd(f, (size_t)n_od * n_ol * 8);
  double* obs_d18o = rd(f, (size_t)n_od * n_ol * 8);
  fclose(f);

  const int n_shards = argc - 3;
  int32_t dev[64];
  for (int s = 0; s < n_shards && s < 64; s++) dev[s] = atoi(argv[3 + s]);
  lwfb_opts o;
  lwfb_default_opts(&o);
  o.n_layers = N; o.simulate_isotopes = iso; o.t0_day = t0; o.n_days = nd; o.record_daily = 0;
  lwfb_multi* mh = NULL;
  CK(lwfb_multi_create(&mh, &o, n_shards, dev));
  CK(lwfb_multi_set_sites(mh, 1, n_rows, forcing));
  CK(lwfb_multi_set_members(mh, M, site, scal, lay, u0, cost));      /* cost-balanced deal */
  CK(lwfb_multi_set_observations(mh, n_od, obs_day, n_ol, obs_layer, obs_theta, obs_d18o));
  CK(lwfb_multi_run(mh));
  double* fin = malloc((size_t)M * L * 8);
  double* scores = malloc((size_t)M * LWFB_NSCORES * 8);
  int32_t* status = malloc((size_t)M * 4 + 8);
  int64_t* counters = malloc((size_t)M * LWFB_NCOUNTERS * 8);
  int32_t* shard = malloc((size_t)M * 4 + 8);
  CK(lwfb_multi_allgather_scores(mh, 0, NULL, 0, NULL, NULL, NULL, NULL, scores));
  CK(lwfb_multi_get_status(mh, status, counters, shard));
  CK(lwfb_multi_get_final(mh, fin));
  float ms_max = 0;
  CK(lwfb_multi_last_run_ms(mh, NULL, &ms_max));
  lwfb_multi_destroy(mh);
  FILE* g = fopen(argv[2], "wb");
  if (!g) return 1;
  fwrite(fin, 8, (size_t)M * L, g);
  fwrite(scores, 8, (size_t)M * LWFB_NSCORES, g);
  fwrite(counters, 8, (size_t)M * LWFB_NCOUNTERS, g);
  fwrite(status, 4, (size_t)M, g);
  fwrite(shard, 4, (size_t)M, g);
  fclose(g);
  printf("multi_host: %lld members on %d shard(s), %.1f ms\n", (long long)M, n_shards, ms_max);
  return 0;
}
