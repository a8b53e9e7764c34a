// lwfb90_tsit5.cuh — Tsit5 tableau in the plain (unrolled) form used by the thread-per-member and the wide kernel.
#pragma once
#include "lwfb90_device.cuh"

namespace lwfb {

// Tsit5 tableau (Tsitouras 2011) as used by OrdinaryDiffEq's Tsit5
__device__ constexpr double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
__device__ constexpr double a21 = 0.161, a31 = -0.008480655492356989, a32 = 0.335480655492357,
                            a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815,
                            a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                            a54 = -0.09249506636175525, a61 = 5.86145544294642, a62 = -12.92096931784711,
                            a63 = 8.159367898576159, a64 = -0.071584973281401, a65 = -0.028269050394068383,
                            a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996,
                            a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
__device__ constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469,
                            bt3 = 0.007880878010261995, bt4 = -0.1447110071732629, bt5 = 0.5823571654525552,
                            bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;

// a(s+1, l+1) for the stage loop (row s = stage s+1) and the stage times
__device__ constexpr double kA[6][5] = {{0, 0, 0, 0, 0},
                                        {a21, 0, 0, 0, 0},
                                        {a31, a32, 0, 0, 0},
                                        {a41, a42, a43, 0, 0},
                                        {a51, a52, a53, a54, 0},
                                        {a61, a62, a63, a64, a65}};
__device__ constexpr double kC[6] = {0, c1, c2, c3, c4, 1.0};

// position of the 8 integrated accumulators inside `accum`
__device__ constexpr int kAccPos[8] = {LWFB_A_FLOW, LWFB_A_SEEP, LWFB_A_SRFL, LWFB_A_SLFL,
                                       LWFB_A_BYFL, LWFB_A_DSFL, LWFB_A_GWFL, LWFB_A_VRFLN};

}  // namespace lwfb
