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
#include <string.h>

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
  if (b.stretch_count) { // tell the host when this warp's stretch of beams has landed (the listener expands intensities meanwhile)
    // every lane of the warp that has a beam meets here (not "whoever happens to be converged": the lanes the report speaks
    // for must all have stored)
    const uint64_t left = b.beam_end - (beam - (threadIdx.x & 31u));
    __syncwarp(left >= 32 ? 0xffffffffu : (1u << (unsigned)left) - 1u);
    if ((threadIdx.x & 31u) == 0u) {
      const uint64_t k = (beam - b.beam_begin) >> b.stretch_shift;
      const uint64_t in_stretch = min((uint64_t)1 << b.stretch_shift, b.beam_end - b.beam_begin - (k << b.stretch_shift));
      // the warp's results, then its count: device scope here (the counter lives on the device); the warp that completes the
      // stretch fences system-wide behind every count it has seen, which orders all the stretch's results before the flag
      // for the host as well (a system-wide fence per warp would make each wait out its own stores' way across the link:
      // +0.05 ms on a 1 M-beam march)
      __threadfence();
      if (atomicAdd(b.stretch_count + k, 1u) + 1u == (uint32_t)((in_stretch + 31) >> 5)) { // the stretch's last warp
        b.stretch_count[k] = 0u; // for the next launch
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t *>(b.stretch_done + k) = b.done_seq;
      }
    }
  }
}

// ---- the march with lane refill ---------------------------------------------------------------------------------
// In k2_ray_march a warp is busy until the longest of its 32 rays has ended: 20.6 of 32 lanes work on average on the
// config-3 world (profiles/). Here a warp owns a CHUNK of consecutive beams and hands them to its lanes one by one: a
// lane whose ray has ended takes the chunk's next beam at the following refill point, so the warp's lanes stay busy until
// the chunk runs dry. What makes that pay:
//  * a ray is a resumable state (trace::RayState) advanced one region step at a time (trace::ray_step);
//  * the expensive part of a ray's set-up, cos and sin of its heading as the host C library gives them, is done by a
//    kernel of its own with every lane busy (k2_beam_trig), so a refill is ~60 instructions and is only done when at
//    least kRefillLanes lanes are free (or nothing else is left to do);
//  * results go through shared memory and leave the chunk together, 256 contiguous bytes per warp store, so that the
//    stores into page-locked host memory (the e2e path) stay whole PCIe writes whatever order the rays ended in.
// Same rays, same arithmetic: the results are those of k2_ray_march bit for bit (tests/test_gpu_synthetic.py).
#ifndef STG_MARCH_CHUNK
#define STG_MARCH_CHUNK 256
#endif
#ifndef STG_MARCH_REFILL_CTAS
#define STG_MARCH_REFILL_CTAS 9 // resident CTAs of 2 warps per SM the register budget is cut for
#endif
constexpr uint32_t kChunkBeams = STG_MARCH_CHUNK; // per warp and chunk: 8 rays per lane on average
#ifndef STG_MARCH_REFILL_LANES
#define STG_MARCH_REFILL_LANES 1 // 1: every lane refills on its own (free-running); > 1: lock-step refill once that many lanes are free
#endif

__global__ void __launch_bounds__(128) k2_beam_trig(FanBatchView b, double *trig)
{
  const uint64_t beam = b.beam_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (beam >= b.beam_end)
    return;
  double cosa, sina;
  trace::heading_trig(b.heading[beam], cosa, sina);
  reinterpret_cast<double2 *>(trig)[beam] = make_double2(cosa, sina);
}

__device__ __forceinline__ uint32_t fan_of_beam(const FanBatchView &b, uint64_t beam)
{
  uint32_t lo = 0, hi = b.n_fans;
  if (b.uniform_samples) {
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
  return lo;
}

template <int kWarps>
__global__ void __launch_bounds__(32 * kWarps, STG_MARCH_REFILL_CTAS) k2_ray_march_refill(GridView g, FanBatchView b, const double *trig, uint32_t n_chunks,
                                                                                         uint32_t *next_chunk)
{
  __shared__ double s_range[kWarps][kChunkBeams];
  __shared__ int32_t s_hit[kWarps][kChunkBeams];
  __shared__ int2 s_cell[kWarps][kChunkBeams];
  __shared__ uint32_t s_next[kWarps];
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // the warps take chunks off one counter: whoever is done first takes the next (chunks differ a lot in length - open
  // space against a cluttered corner - so a fixed share per warp leaves the device half empty towards the end)
  for (;;) {
    uint32_t chunk = 0;
    if (lane == 0)
      chunk = atomicAdd(next_chunk, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= n_chunks)
      break;
    const uint64_t first = b.beam_begin + (uint64_t)chunk * kChunkBeams;
    const uint32_t chunk_n = (uint32_t)min((uint64_t)kChunkBeams, b.beam_end - first);
    if (lane == 0)
      s_next[wib] = 0;
    __syncwarp();
#if STG_MARCH_REFILL_LANES > 1
    // lock-step variant (measured slower, kept for the record): refill when enough lanes are free, all lanes meet every step
    uint32_t next = 0;
    bool busy = false;
    uint32_t slot = 0;
    trace::RayState ray;
    for (;;) {
      const uint32_t idle = __ballot_sync(0xffffffffu, !busy);
      if (idle == 0xffffffffu && next >= chunk_n)
        break;
      if (next < chunk_n && ((uint32_t)__popc(idle) >= STG_MARCH_REFILL_LANES || idle == 0xffffffffu)) {
        const uint32_t mine = next + (uint32_t)__popc(idle & lt_mask);
        if (!busy && mine < chunk_n) {
          slot = mine;
          const uint64_t beam = first + mine;
          const FanRec *fan = b.fans + fan_of_beam(b, beam);
          const double2 cs = __ldg(reinterpret_cast<const double2 *>(trig) + beam);
          trace::ray_setup(g, fan, cs.x, cs.y, ray);
          busy = true;
          if (ray.n <= 0) {
            s_range[wib][slot] = fan->range;
            s_hit[wib][slot] = -1;
            s_cell[wib][slot] = make_int2(0, 0);
            busy = false;
          }
        }
        next = min(next + (uint32_t)__popc(idle), chunk_n);
      }
      if (busy) {
        int32_t hit;
        if (trace::ray_step(g, ray, hit)) {
          s_range[wib][slot] = trace::ray_range(g, ray, hit);
          s_hit[wib][slot] = hit;
          s_cell[wib][slot] = hit >= 0 ? make_int2(ray.hit_x, ray.hit_y) : make_int2(0, 0);
          busy = false;
        }
      }
    }
#else
    // Every lane is a worker of its own: it takes the chunk's next beam off a counter in shared memory the moment its ray
    // has ended, without meeting the other lanes - lanes at the same point of the walk (of whatever ray) still execute
    // together, as in k2_ray_march, and none waits for a neighbour's longer region.
    (void)lt_mask;
    for (;;) {
      const uint32_t slot = atomicAdd(&s_next[wib], 1u);
      if (slot >= chunk_n)
        break;
      const uint64_t beam = first + slot;
      const FanRec *fan = b.fans + fan_of_beam(b, beam);
      const double2 cs = __ldg(reinterpret_cast<const double2 *>(trig) + beam);
      trace::RayState ray;
      trace::ray_setup(g, fan, cs.x, cs.y, ray);
      int32_t hit = -1;
      while (ray.n > 0 && !trace::ray_step(g, ray, hit)) {
      }
      s_range[wib][slot] = trace::ray_range(g, ray, hit);
      s_hit[wib][slot] = hit;
      s_cell[wib][slot] = hit >= 0 ? make_int2(ray.hit_x, ray.hit_y) : make_int2(0, 0);
    }
#endif
    __syncwarp();
    // the chunk leaves together: K3's ranger epilogue (model_ranger.cc:243-251), coalesced
    for (uint32_t i = lane; i < chunk_n; i += 32) {
      const uint64_t beam = first + i;
      double range = s_range[wib][i];
      const int32_t hit = s_hit[wib][i];
      if (b.noise) { // "Apply noise only if it is in valid range" (:244-248)
        const uint32_t f = fan_of_beam(b, beam);
        const FanNoiseRec nz = b.noise[f];
        const FanRec *fan = b.fans + f;
        if ((nz.range_noise != 0.0 || nz.range_noise_const != 0.0) && range < fan->range) {
          const unsigned long long seed = nz.seed ? nz.seed : fan->finder + 1ull;
          const double u = noise_simple(noise_bits(seed, beam, b.noise_epoch, 1));
          double r1 = (double)(noise_bits(seed, beam, b.noise_epoch, 2) >> 11) * (1.0 / 9007199254740992.0);
          const double r2 = (double)(noise_bits(seed, beam, b.noise_epoch, 3) >> 11) * (1.0 / 9007199254740992.0) * 6.283185307179586;
          if (r1 < 1e-100)
            r1 = 1e-100;
          const double gauss = sqrt(nz.range_noise_const * (-2.0 * log(r1))) * cos(r2);
          range = range + range * nz.range_noise * u + gauss;
        }
      }
      if (b.ranges)
        b.ranges[beam] = range;
      if (b.intensities)
        b.intensities[beam] = hit >= 0 ? __ldg(g.mdl_ranger_return + hit) : 0.0;
      if (b.intens_idx)
        b.intens_idx[beam] = hit >= 0 ? __ldg(g.mdl_pal + hit) : (uint8_t)0;
      if (b.hit_model)
        b.hit_model[beam] = hit;
      if (b.hit_cell)
        reinterpret_cast<int2 *>(b.hit_cell)[beam] = s_cell[wib][i];
    }
    __syncwarp();
  }
}

} // namespace

void launch_beam_trig(const FanBatchView &b, double *trig, void *stream)
{
  if (b.beam_end <= b.beam_begin)
    return;
  const unsigned blocks = (unsigned)((b.beam_end - b.beam_begin + 127) / 128);
  k2_beam_trig<<<blocks, 128, 0, (cudaStream_t)stream>>>(b, trig);
}

// The refill march needs per-beam (cos, sin) in `trig` (strict mode's host values, or k2_beam_trig's) and serves
// everything but the step counters, which stay with k2_ray_march.
bool ray_march_refill_wanted(const FanBatchView &b)
{
  // Off unless asked for (STG_RT_MARCH=refill): on the config-3 world it is SLOWER than k2_ray_march, 0.355 ms against
  // 0.177 ms (profiles/r02_march_refill.md) - adjacent beams march through the same tiles at the same time, and that
  // coherence is worth more than the idle lanes it leaves; handing lanes other beams breaks it.
  const char *e = getenv("STG_RT_MARCH");
  return e && !strcmp(e, "refill") && !b.steps && b.n_beams >= 4096 && b.n_beams >= 32ull * b.n_fans;
}

void launch_ray_march_refill(const GridView &g, const FanBatchView &b, const double *trig, uint32_t *work_counter, void *stream)
{
  if (b.beam_end <= b.beam_begin)
    return;
  const uint64_t n = b.beam_end - b.beam_begin;
  const uint32_t n_chunks = (uint32_t)((n + kChunkBeams - 1) / kChunkBeams);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  constexpr int kWarps = 2;
  const uint32_t resident = (uint32_t)sms * STG_MARCH_REFILL_CTAS;    // CTAs the device holds at once: one wave, persistent
  const uint32_t blocks = min((n_chunks + kWarps - 1) / kWarps, resident);
  cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), (cudaStream_t)stream);
  k2_ray_march_refill<kWarps><<<blocks, 32 * kWarps, 0, (cudaStream_t)stream>>>(g, b, trig, n_chunks, work_counter);
}

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
