// Closed forms of the reference's cell-by-cell region walk (world.cc:823-882), shared by the march (rt_trace.cuh)
// and, compiled for the host, by tests/test_walk_closed_form.py, which holds every function here to the step-by-step
// loop over exhaustive small cases and random large ones.
//
// The walk in run form (rt_trace.cuh): the ray stands in cell (ap, cp) of a 32 x 32 region, both coordinates counted
// in the direction of travel, with error term E:   E < 0 -> run step:   ap += 1, E += b_run
//                                                   else  -> cross step: cp += 1, E -= b_cross
// (b_cross >= b_run >= 0, b_cross > 0). After i run steps and j cross steps E = E0 + i*b_run - j*b_cross, and
//   * the j-th cross step (j >= 1) is taken after exactly  i_j = max(0, ceil(((j-1)*b_cross - E0) / b_run))  run steps,
//   * the i-th run step  (i >= 1) is taken after exactly  j_i = E0 + (i-1)*b_run < 0 ? 0 : floor((E0 + (i-1)*b_run) / b_cross) + 1
//     cross steps,
// so where the walk leaves the region, and in which state, is one division away. No rounding is involved: integers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define STG_HD __host__ __device__ __forceinline__
#else
#define STG_HD inline
#endif

namespace stg {
namespace walk {

constexpr int32_t kWidth = 32; // kRegionWidth (rt_device.h)

// The closed forms multiply b_cross by up to 33 in 32-bit arithmetic: rays longer than this many cells along their
// major axis (335 km at 2 cm) take the step-by-step walk instead.
constexpr int32_t kMaxMajorCells = 1 << 24;

// floor(x / b) for 0 <= x, 0 < b, quotient <= 64: a float estimate (relative error < 2^-21, so off by at most one)
// and one correction either way.
STG_HD int32_t floor_div_small(int32_t x, int32_t b)
{
#if defined(__CUDA_ARCH__)
  float rb;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"((float)b));
  int32_t q = __float2int_rz((float)x * rb);
#else
  int32_t q = (int32_t)((float)x * (1.0f / (float)b));
#endif
  q -= (q * b > x) ? 1 : 0;
  q += ((q + 1) * b <= x) ? 1 : 0;
  return q;
}

// The state in which the walk from (ap0, cp0, E0) leaves the region when nothing stops it and the ray does not end
// inside: ap == 32 (through the run side) or cp == 32 (through the cross side), E as the loop leaves it. The number of
// cells visited is (ap - ap0) + (cp - cp0).
STG_HD void region_exit(int32_t ap0, int32_t cp0, int32_t E0, int32_t b_run, int32_t b_cross, int32_t &ap, int32_t &cp, int32_t &E)
{
  const int32_t a1 = kWidth - 1 - ap0;   // run steps that stay inside
  const int32_t c1 = kWidth - 1 - cp0;   // cross steps that stay inside
  const int32_t need = c1 * b_cross - E0; // the cross step that leaves is taken once i * b_run >= need
  const int32_t have = a1 * b_run;
  if (need <= have) { // ... which happens while ap is still inside: out through the cross side
    const int32_t i = need > 0 ? floor_div_small(need + b_run - 1, b_run) : 0; // need > 0 here implies b_run > 0
    ap = ap0 + i;
    cp = kWidth;
    E = E0 + i * b_run - (c1 + 1) * b_cross;
  } else { // out through the run side, by run step number a1 + 1
    const int32_t t = E0 + have;
    const int32_t j = t >= 0 ? floor_div_small(t, b_cross) + 1 : 0;
    ap = kWidth;
    cp = cp0 + j;
    E = E0 + (a1 + 1) * b_run - j * b_cross;
  }
}

// The state in which the walk first stands in cross coordinate cp0 + j (1 <= j, a coordinate the walk reaches inside
// the region): right after its j-th cross step.
STG_HD void jump_cross(int32_t ap0, int32_t cp0, int32_t E0, int32_t b_run, int32_t b_cross, int32_t j, int32_t &ap, int32_t &cp, int32_t &E)
{
  const int32_t need = (j - 1) * b_cross - E0;
  const int32_t i = need > 0 ? floor_div_small(need + b_run - 1, b_run) : 0;
  ap = ap0 + i;
  cp = cp0 + j;
  E = E0 + i * b_run - j * b_cross;
}

// The state in which the walk first stands in run coordinate ap0 + i (1 <= i, reached inside the region): right after
// its i-th run step.
STG_HD void jump_run(int32_t ap0, int32_t cp0, int32_t E0, int32_t b_run, int32_t b_cross, int32_t i, int32_t &ap, int32_t &cp, int32_t &E)
{
  const int32_t t = E0 + (i - 1) * b_run;
  const int32_t j = t >= 0 ? floor_div_small(t, b_cross) + 1 : 0;
  ap = ap0 + i;
  cp = cp0 + j;
  E = E0 + i * b_run - j * b_cross;
}

} // namespace walk
} // namespace stg
