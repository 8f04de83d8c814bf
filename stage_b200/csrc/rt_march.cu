// K2: the batched ray march (World::Raytrace, world.cc:756-963) with the ranger epilogue K3 fused
// (model_ranger.cc:243-251). One lane per beam; consecutive lanes are consecutive beams of a fan, so
// the lanes of a warp read the same few 256-byte occupancy tiles and stop at about the same wall.
//
// The result is the reference's, bit for bit, but the walk is organised for the GPU:
//  * Inside a populated region the reference steps cell by cell on a double position. Here the
//    region walk is pure integer work on the region's occupancy masks, and the position is brought
//    up to date once, at the region exit or at the hit (seq_add below reproduces the rounding of
//    the reference's repeated `glob += 1.0`).
//  * The cells of one row that the line visits between two y steps form a run; a run is tested with
//    ONE AND of the row mask (x-major rays) or column mask (y-major rays) instead of one load and
//    test per cell. Only where a mask bit is set is the cell's content looked up (two dependent
//    loads: cell-entry slot, then the 32-byte BlockRec).
//  * Empty regions are skipped with the reference's own crossing arithmetic (FP64, same operation
//    order, no FMA contraction: build with -fmad=false), including its stale error term, the
//    truncating double->int conversions and the int -= double of the remaining distance.
#include <stdlib.h>

#include "rt_trace.cuh"

namespace stg {

namespace {

#ifndef STG_MARCH_CTAS_PER_SM
#define STG_MARCH_CTAS_PER_SM 9 // of 128 threads: 56 registers per thread (measured: 0.179 ms; 8 = 64 regs 0.184, 10 = 48 regs 0.193, 7 = 72 regs slower still)
#endif

template <bool kSteps>
__global__ void __launch_bounds__(128, STG_MARCH_CTAS_PER_SM) k2_ray_march(GridView g, FanBatchView b)
{
  const uint64_t beam = b.beam_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (beam >= b.beam_end)
    return;

  // which fan does this beam belong to
  uint32_t lo = 0, hi = b.n_fans;
  if (b.uniform_samples) { // the usual case: one sensor type
    lo = (uint32_t)beam / b.uniform_samples;
    hi = lo + 1;
  }
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(b.fan_first_beam + mid) <= beam)
      lo = mid;
    else
      hi = mid;
  }
  const FanRec *fan = b.fans + lo;

  double cosa, sina;
  if (b.trig_in) { // strict mode: the host libm's values
    cosa = b.trig_in[2 * beam];
    sina = b.trig_in[2 * beam + 1];
  } else {
    trace::heading_trig(b.heading[beam], cosa, sina);
  }
  trace::RayResult r;
  trace::trace_ray<kSteps>(g, fan, cosa, sina, r);

  // K3 (ranger epilogue, fused): model_ranger.cc:243-251
  if (b.noise) { // "Apply noise only if it is in valid range" (:244-248)
    const FanNoiseRec nz = b.noise[lo];
    if ((nz.range_noise != 0.0 || nz.range_noise_const != 0.0) && r.range < fan->range) {
      const unsigned long long seed = nz.seed ? nz.seed : fan->finder + 1ull;
      const double u = noise_simple(noise_bits(seed, beam, b.noise_epoch, 1));
      // generateGaussianNoise(variance), :181-200 (Box-Muller; the reference alternates cos / sin of one
      // pair of draws, every beam here takes the cos branch of its own pair: same N(0, variance))
      double r1 = (double)(noise_bits(seed, beam, b.noise_epoch, 2) >> 11) * (1.0 / 9007199254740992.0);
      const double r2 = (double)(noise_bits(seed, beam, b.noise_epoch, 3) >> 11) * (1.0 / 9007199254740992.0) * 6.283185307179586;
      if (r1 < 1e-100)
        r1 = 1e-100;
      const double gauss = sqrt(nz.range_noise_const * (-2.0 * log(r1))) * cos(r2);
      r.range = r.range + r.range * nz.range_noise * u + gauss;
    }
  }
  if (b.ranges)
    b.ranges[beam] = r.range;
  if (b.intensities)
    b.intensities[beam] = r.hit_model >= 0 ? __ldg(g.mdl_ranger_return + r.hit_model) : 0.0;
  if (b.intens_idx)
    b.intens_idx[beam] = r.hit_model >= 0 ? __ldg(g.mdl_pal + r.hit_model) : (uint8_t)0;
  if (b.hit_model)
    b.hit_model[beam] = r.hit_model;
  if (b.hit_cell) {
    b.hit_cell[2 * beam] = r.hit_x;
    b.hit_cell[2 * beam + 1] = r.hit_y;
  }
  if (kSteps) {
    b.steps[2 * beam] = r.cell_visits;
    b.steps[2 * beam + 1] = r.region_probes;
  }
}

} // namespace

void launch_ray_march(const GridView &g, const FanBatchView &b, void *stream)
{
  if (b.beam_end <= b.beam_begin)
    return;
  static int threads = 0;
  if (!threads) { // experiment knob; the default is what the profiles under profiles/ were taken with
    const char *e = getenv("STG_RT_MARCH_BLOCK");
    threads = e ? atoi(e) : 64;
    if (threads != 32 && threads != 64 && threads != 128)
      threads = 64;
  }
  const unsigned blocks = (unsigned)((b.beam_end - b.beam_begin + threads - 1) / threads);
  if (b.steps)
    k2_ray_march<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(g, b);
  else
    k2_ray_march<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(g, b);
}

} // namespace stg
