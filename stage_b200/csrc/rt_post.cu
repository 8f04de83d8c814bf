// K3: sensor post-processing kernels that sit on top of the ray tracer.
//   k3_fiducials  ModelFiducial::Update + AddModelIfVisible (model_fiducial.cc:109-293): candidate
//                 culls, one line-of-sight ray per candidate, accept / reduce into Fiducial records
//   k3_blobs      ModelBlobfinder::Update (model_blobfinder.cc:165-250): run-length colour
//                 segmentation of a traced scan line into Blob records
// (the ranger's epilogue - intensity, bearing - is fused into the march kernel, rt_march.cu)
#include "rt_trace.cuh"

namespace stg {

namespace {

// normalize(), stage.hh:163-170
__device__ __forceinline__ double normalize_angle(double a)
{
  while (a < -M_PI)
    a += 2.0 * M_PI;
  while (a > M_PI)
    a -= 2.0 * M_PI;
  return a;
}

// AddModelIfVisible (model_fiducial.cc:109-224) for one (finder, target) pair, in three parts so that a warp can first
// sift a finder's candidates with all lanes and then trace only the survivors, again with all lanes (k3_fiducials_grid):
//   fid_cull    the x / y window of the candidate query (:245-262, bounds inclusive), key, range, field of view, relatedness
//   fid_sees    the line-of-sight ray and the accept rule (:179-193)
//   fid_record  the Fiducial record (:199-221)
__device__ __forceinline__ bool fid_cull(const GridView &g, const FidFinderRec &fd, uint32_t finder_root, const FidTargetRec &him, double &range,
                                         double &dtheta)
{
  if (him.model >= g.n_models) // a padding record of a rank's slot in the gathered target array, or an id nobody published
    return false;
  if (him.model == fd.finder || him.x < fd.x - fd.max_range_anon || him.x > fd.x + fd.max_range_anon ||
      him.y < fd.y - fd.max_range_anon || him.y > fd.y + fd.max_range_anon)
    return false;
  if (him.fiducial_key != fd.fiducial_key) // :117
    return false;
  const double dx = him.x - fd.x, dy = him.y - fd.y;
  range = hypot(dy, dx);                                            // :130
  if (range >= fd.max_range_anon)                                   // :134
    return false;
  dtheta = normalize_angle(atan2(dy, dx) - fd.a);                  // :144-145
  if (fabs(dtheta) > fd.fov / 2.0)                                  // :147
    return false;
  return __ldg(g.mdl_root + him.model) != finder_root;             // IsRelated, :152
}

__device__ __forceinline__ bool fid_sees(const GridView &g, const FidFinderRec &fd, const FanRec *fan, uint32_t him_model, double dtheta)
{
  // Raytrace(Pose(0,0,0,dtheta), max_range_anon, fiducial_raytrace_match, NULL, true), :179:
  // LocalToGlobal adds the heading onto (GetGlobalPose() + geom.pose) and normalises it
  double cosa, sina;
  trace::heading_trig(normalize_angle(fd.oa + dtheta), cosa, sina);
  trace::RayResult r;
  trace::trace_ray<false>(g, fan, cosa, sina, r);
  int32_t seen = r.hit_model;
  if (fd.ignore_zloc && seen < 0) // :184
    seen = (int32_t)him_model;
  return seen == (int32_t)him_model; // :192
}

__device__ __forceinline__ void fid_record(const FidFinderRec &fd, const FidTargetRec &him, uint32_t t, double range, double dtheta, FiducialRec &rec)
{
  rec.target = t;
  rec.id = range < fd.max_range_id ? him.fiducial_return : 0; // :219
  rec.range = range;
  rec.bearing = dtheta;
  rec.geom[0] = him.sx;
  rec.geom[1] = him.sy;
  rec.geom[2] = him.sz;
  rec.geom[3] = normalize_angle(him.a - fd.a); // :209
  rec.pose[0] = him.x;
  rec.pose[1] = him.y;
  rec.pose[2] = him.z;
  rec.pose[3] = him.a;
}

__device__ __forceinline__ bool fid_try(const GridView &g, const FidFinderRec &fd, const FanRec *fan, uint32_t finder_root,
                                        const FidTargetRec &him, uint32_t t, FiducialRec &rec)
{
  double range, dtheta;
  if (!fid_cull(g, fd, finder_root, him, range, dtheta) || !fid_sees(g, fd, fan, him.model, dtheta))
    return false;
  fid_record(fd, him, t, range, dtheta, rec);
  return true;
}

// One warp per finder. Lanes take 32 targets at a time, in the caller's order (the reference visits
// candidates in ascending Model* order, model_fiducial.cc:266-283, and controllers depend on it), so
// the accepted records are compacted with a ballot prefix to keep that order.
__global__ void __launch_bounds__(128) k3_fiducials(GridView g, FidBatchView b)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= b.n_finders)
    return;
  const FidFinderRec fd = b.finders[warp];
  const FanRec *fan = b.fans + warp; // the line-of-sight ray's constants (origin, range, predicate)
  const uint32_t finder_root = __ldg(g.mdl_root + fd.finder);
  uint32_t n_out = 0;
  for (uint32_t t0 = 0; t0 < b.n_targets; t0 += 32) {
    const uint32_t t = t0 + lane;
    FiducialRec rec;
    const bool accept = t < b.n_targets && fid_try(g, fd, fan, finder_root, b.targets[t], t, rec);
    const uint32_t votes = __ballot_sync(0xffffffffu, accept);
    if (accept) {
      const uint32_t slot = n_out + __popc(votes & ((1u << lane) - 1u));
      if (slot < b.max_per_finder)
        b.out[(size_t)warp * b.max_per_finder + slot] = rec;
    }
    n_out += __popc(votes);
  }
  if (lane == 0)
    b.out_count[warp] = n_out; // may exceed max_per_finder: the caller sees the truncation
}

// ---- the same with a uniform-grid candidate query -----------------------------
// The reference rebuilds two position-sorted std::sets every tick (world.cc:623-629) and intersects
// two range queries per finder (model_fiducial.cc:245-283). Here the targets are binned into square
// cells at least as wide as the longest max_range_anon, so a finder's +-range window touches at
// most 3 x 3 bins; bins are filled by a counting sort (count, scan, scatter) every submit.
__global__ void __launch_bounds__(256) k3_fid_bin_count(FidGridView v, const FidTargetRec *targets, uint32_t n_targets)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_targets)
    return;
  const int32_t bx = min(max((int32_t)floor((targets[t].x - v.x0) / v.cell), 0), v.nbx - 1);
  const int32_t by = min(max((int32_t)floor((targets[t].y - v.y0) / v.cell), 0), v.nby - 1);
  const uint32_t bin = (uint32_t)(by * v.nbx + bx);
  v.bin_of[t] = bin;
  atomicAdd(v.start + bin, 1u);
}

// exclusive scan of v.start[0..n_bins) in place, start[n_bins] = total; one CTA, chunk per thread
__global__ void __launch_bounds__(1024) k3_fid_bin_scan(FidGridView v)
{
  __shared__ uint32_t partial[1024];
  const uint32_t n = (uint32_t)(v.nbx * v.nby), tid = threadIdx.x;
  const uint32_t per = (n + 1023) / 1024, lo = min(tid * per, n), hi = min(lo + per, n);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++)
    sum += v.start[i];
  partial[tid] = sum;
  __syncthreads();
  for (uint32_t o = 1; o < 1024; o <<= 1) { // Hillis-Steele inclusive scan of the partial sums
    const uint32_t add = tid >= o ? partial[tid - o] : 0;
    __syncthreads();
    partial[tid] += add;
    __syncthreads();
  }
  uint32_t run = tid ? partial[tid - 1] : 0;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t c = v.start[i];
    v.start[i] = run;
    v.cursor[i] = run;
    run += c;
  }
  if (tid == 1023)
    v.start[n] = partial[1023];
}

__global__ void __launch_bounds__(256) k3_fid_bin_scatter(FidGridView v, uint32_t n_targets)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_targets)
    v.sorted[atomicAdd(v.cursor + v.bin_of[t], 1u)] = t;
}

// One warp per finder, in two phases so that both run with full lanes. The lanes SIFT the candidates of the window's bins 32
// at a time (fid_cull: cheap, and four of five fall out - 2100 in the 3 x 3 bins of a config-4 finder, 370 inside its half
// disc) and collect the survivors in shared memory; then the warp SORTS what it has collected by bearing (a counting sort over
// kFidSectors sectors of the field of view) and TRACES the lines of sight 32 at a time in that order (fid_sees). All of a
// finder's rays leave from one point, so neighbours in bearing cross the same cells and - in a swarm, where almost every line
// of sight ends on the nearest body in its direction (1.7 % of config 4's candidates are seen) - end on the same body after the
// same number of steps: a warp's rays then march as one. Traced in the order the bins delivered them they shared nothing:
// 14.7 of 32 lanes busy, 13.1 ms for config 4's 37 M rays (profiles/). Tracing straight from the sift: 9 of 32 lanes, 16.1 ms.
constexpr uint32_t kFidCap = 512;     // survivors a warp collects before it sorts and traces them (more: in rounds)
constexpr uint32_t kFidSectors = 128; // bearing classes of the counting sort
#ifndef STG_FID_MIN_BLOCKS
#define STG_FID_MIN_BLOCKS 4
#endif

__global__ void __launch_bounds__(128, STG_FID_MIN_BLOCKS) k3_fiducials_grid(GridView g, FidBatchView b, FidGridView v)
{
  __shared__ uint32_t s_t[4][kFidCap];
  __shared__ double s_dtheta[4][kFidCap];
  __shared__ uint16_t s_order[4][kFidCap];
  __shared__ uint32_t s_sector[4][kFidSectors];
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (warp >= b.n_finders)
    return;
  const FidFinderRec fd = b.finders[warp];
  const FanRec *fan = b.fans + warp;
  const uint32_t finder_root = __ldg(g.mdl_root + fd.finder);
  const uint32_t lt = (1u << lane) - 1u;
  FiducialRec *out = b.out + (size_t)warp * b.max_per_finder;
  const int32_t bx0 = min(max((int32_t)floor((fd.x - fd.max_range_anon - v.x0) / v.cell), 0), v.nbx - 1);
  const int32_t bx1 = min(max((int32_t)floor((fd.x + fd.max_range_anon - v.x0) / v.cell), 0), v.nbx - 1);
  const int32_t by0 = min(max((int32_t)floor((fd.y - fd.max_range_anon - v.y0) / v.cell), 0), v.nby - 1);
  const int32_t by1 = min(max((int32_t)floor((fd.y + fd.max_range_anon - v.y0) / v.cell), 0), v.nby - 1);
  const float sector_scale = fd.fov > 0 ? (float)((double)kFidSectors / fd.fov) : 0.0f, half_fov = (float)(fd.fov / 2.0);
  uint32_t n_out = 0, n_kept = 0;
  // sort the n_kept collected survivors by bearing, trace them in that order, keep the records of those that are seen
  auto trace_kept = [&]() {
    auto sector_of = [&](uint32_t k) {
      return (uint32_t)min(max((int32_t)(((float)s_dtheta[wib][k] + half_fov) * sector_scale), 0), (int32_t)kFidSectors - 1);
    };
    for (uint32_t i = lane; i < kFidSectors; i += 32)
      s_sector[wib][i] = 0;
    __syncwarp();
    for (uint32_t k = lane; k < n_kept; k += 32)
      atomicAdd(&s_sector[wib][sector_of(k)], 1u);
    __syncwarp();
    { // exclusive scan of the sector counts: kFidSectors / 32 consecutive sectors per lane, a warp scan of the lanes' sums
      constexpr uint32_t kPer = kFidSectors / 32;
      uint32_t c[kPer], sum = 0;
      for (uint32_t j = 0; j < kPer; j++) {
        c[j] = s_sector[wib][lane * kPer + j];
        sum += c[j];
      }
      uint32_t incl = sum;
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d)
          incl += up;
      }
      uint32_t run = incl - sum;
      for (uint32_t j = 0; j < kPer; j++) {
        s_sector[wib][lane * kPer + j] = run;
        run += c[j];
      }
    }
    __syncwarp();
    for (uint32_t k = lane; k < n_kept; k += 32)
      s_order[wib][atomicAdd(&s_sector[wib][sector_of(k)], 1u)] = (uint16_t)k;
    __syncwarp();
    for (uint32_t i0 = 0; i0 < n_kept; i0 += 32) {
      bool accept = false;
      uint32_t t = 0;
      double dtheta = 0;
      if (i0 + lane < n_kept) {
        const uint32_t k = s_order[wib][i0 + lane];
        t = s_t[wib][k];
        dtheta = s_dtheta[wib][k];
        accept = fid_sees(g, fd, fan, __ldg(&b.targets[t].model), dtheta);
      }
      const uint32_t votes = __ballot_sync(0xffffffffu, accept);
      if (accept) {
        const uint32_t slot = n_out + __popc(votes & lt);
        if (slot < b.max_per_finder) {
          const FidTargetRec &him = b.targets[t];
          fid_record(fd, him, t, hypot(him.y - fd.y, him.x - fd.x), dtheta, out[slot]); // the range fid_cull computed (:130)
        }
      }
      n_out += __popc(votes);
    }
    n_kept = 0;
    __syncwarp();
  };
  // The bins are a quarter of the longest detection range wide and a row of bins is one stretch of the sorted target list,
  // so a row is sifted in one go - and only the part of it that can hold a target this finder accepts: the row's band cut
  // to the circle of radius max_range_anon (fid_cull rejects range >= that) and, for a field of view of at most pi, to the
  // half plane in front of the finder (it rejects |bearing| > fov / 2 >= pi / 2 behind it), both taken a little too wide;
  // every exact test stays with fid_cull. (Targets outside the bins' extent sit in the border bins: the cut is made on
  // coordinates, and a coordinate outside the extent maps to the border bin like the target did.)
  const double R = fd.max_range_anon, slack = 1e-6 * (R + 1.0);
  const bool front_only = fd.fov <= M_PI;
  const double hx = cos(fd.a), hy = sin(fd.a);
  for (int32_t by = by0; by <= by1; by++) {
    const double band_lo = by == 0 ? -R : fmax(v.y0 + by * v.cell - fd.y, -R), band_hi = by == v.nby - 1 ? R : fmin(v.y0 + (by + 1) * v.cell - fd.y, R);
    if (band_lo > band_hi + slack)
      continue;
    const double dy_near = band_lo > 0 ? band_lo : band_hi < 0 ? -band_hi : 0.0; // the band's distance from the finder
    const double half = sqrt(fmax(R * R - dy_near * dy_near, 0.0)) + slack;
    double xlo = -half, xhi = half;
    if (front_only) { // dx * hx + dy * hy >= -slack for some dy of the band
      const double need = -slack - fmax(band_lo * hy, band_hi * hy);
      if (hx > 1e-9)
        xlo = fmax(xlo, need / hx);
      else if (hx < -1e-9)
        xhi = fmin(xhi, need / hx);
      else if (need > 0)
        continue;
      if (xlo > xhi)
        continue;
    }
    const int32_t rx0 = min(max((int32_t)floor((fd.x + xlo - v.x0) / v.cell), bx0), bx1), rx1 = min(max((int32_t)floor((fd.x + xhi - v.x0) / v.cell), bx0), bx1);
    const uint32_t lo = __ldg(v.start + (uint32_t)(by * v.nbx + rx0)), hi = __ldg(v.start + (uint32_t)(by * v.nbx + rx1) + 1);
    for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool pass = false;
      uint32_t t = 0;
      double range = 0, dtheta = 0;
      if (i < hi) {
        t = __ldg(v.sorted + i);
        pass = fid_cull(g, fd, finder_root, b.targets[t], range, dtheta);
      }
      const uint32_t votes = __ballot_sync(0xffffffffu, pass);
      if (pass) { // n_kept <= kFidCap - 32 here
        const uint32_t at = n_kept + __popc(votes & lt);
        s_t[wib][at] = t;
        s_dtheta[wib][at] = dtheta;
      }
      n_kept += __popc(votes);
      __syncwarp();
      if (n_kept > kFidCap - 32)
        trace_kept();
    }
  }
  if (n_kept)
    trace_kept();
  __syncwarp();
  if (lane == 0) {
    // candidates were traced in order of bearing: put the records back into target order (ascending Model*)
    const uint32_t k = min(n_out, b.max_per_finder);
    for (uint32_t i = 1; i < k; i++) {
      const FiducialRec r = out[i];
      uint32_t j = i;
      while (j > 0 && out[j - 1].target > r.target) {
        out[j] = out[j - 1];
        j--;
      }
      out[j] = r;
    }
    b.out_count[warp] = n_out; // beyond max_per_finder the kept records are an unspecified subset
  }
}

// ---- packing the fiducial records for the trip home ---------------------------------------------
// out[] is fixed-stride (max_per_finder records per finder) and mostly empty: at BASELINE config 4
// (100k finders) it is 845 MB of which 60 MB are records. Only the records cross PCIe: offsets =
// exclusive scan of min(count, max_per_finder), then one warp per finder copies its records, 8-byte
// word by word, into page-locked host memory (consecutive finders -> consecutive addresses).
__global__ void __launch_bounds__(1024) k3_fid_pack_scan(FidBatchView b, uint32_t *offsets)
{
  __shared__ uint32_t partial[1024];
  const uint32_t n = b.n_finders, tid = threadIdx.x;
  const uint32_t per = (n + 1023) / 1024, lo = min(tid * per, n), hi = min(lo + per, n);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++)
    sum += min(b.out_count[i], b.max_per_finder);
  partial[tid] = sum;
  __syncthreads();
  for (uint32_t o = 1; o < 1024; o <<= 1) {
    const uint32_t add = tid >= o ? partial[tid - o] : 0;
    __syncthreads();
    partial[tid] += add;
    __syncthreads();
  }
  uint32_t run = tid ? partial[tid - 1] : 0;
  for (uint32_t i = lo; i < hi; i++) {
    offsets[i] = run;
    run += min(b.out_count[i], b.max_per_finder);
  }
  if (tid == 1023)
    offsets[n] = partial[1023];
}

__global__ void __launch_bounds__(128) k3_fid_pack(FidBatchView b, const uint32_t *offsets, unsigned long long *packed,
                                                   uint32_t *count_home)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= b.n_finders)
    return;
  constexpr uint32_t kWords = sizeof(FiducialRec) / 8;
  const uint32_t cnt = b.out_count[warp], n_words = min(cnt, b.max_per_finder) * kWords;
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(b.out + (size_t)warp * b.max_per_finder);
  unsigned long long *dst = packed + (size_t)offsets[warp] * kWords;
  for (uint32_t i = lane; i < n_words; i += 32)
    dst[i] = src[i];
  if (lane == 0)
    count_home[warp] = cnt;
}

// ---- K4: collision tests -----------------------------------------------------------------------
// Model::TestCollision (model.cc:666-679) -> BlockGroup::TestCollision (blockgroup.cc:41-50) ->
// Block::TestCollision (block.cc:129-168), for a batch of models. One warp per query. The cells a
// block is rendered into are not stored per block on the device; they are re-derived, in rendering
// order, from its outline (cell visit k of an edge in closed form, as K1 does). Lane i looks at
// visit t0 + i: every block listed in that cell that belongs to another model, is an obstacle, is
// not related and overlaps in z is a candidate, the one first in the cell's list (smallest seq)
// is what this visit would return; the lowest visit with a candidate wins (ballot), which is the
// model the reference's nested loops return first. The host has already dropped blocks whose own
// model is no obstacle and cut the list at a block that reaches below the floor (rt_api.cu).
__device__ __forceinline__ void edge_cell_post(const EdgeRec &e, uint32_t k, int32_t &x, int32_t &y)
{ // same arithmetic as edge_cell in rt_kernels.cu (World::MapPoly's walk, world.cc:1000-1052)
  const int32_t dx = e.x1 - e.x0, dy = e.y1 - e.y0;
  const int32_t sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
  const long long ax = dx < 0 ? -(long long)dx : dx, ay = dy < 0 ? -(long long)dy : dy;
  const long long i = (2 * ax * (long long)k + ax + ay - 1) / (2 * (ax + ay));
  x = e.x0 + sx * (int32_t)i;
  y = e.y0 + sy * (int32_t)((long long)k - i);
}

__global__ void __launch_bounds__(128) k4_collisions(GridView g, CollisionBatchView b)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= b.n_queries)
    return;
  const CollisionQuery q = b.queries[warp];
  const EdgeRec *edges = b.edges + q.first_edge;
  const int L = q.layer & 1;
  const unsigned long long *slots = g.slots[L];
  const uint4 *recs = reinterpret_cast<const uint4 *>(g.blk_rec[L]);
  int32_t result = -1;
  for (uint32_t t0 = 0; t0 < q.n_steps && result < 0; t0 += 32) {
    const uint32_t t = t0 + lane;
    int32_t mine = -1;
    if (t < q.n_steps) {
      uint32_t ei = 0;
      while (ei + 1 < q.n_edges && edges[ei + 1].first_step <= t)
        ei++;
      const EdgeRec e = edges[ei];
      int32_t x, y;
      edge_cell_post(e, t - e.first_step, x, y);
      const uint4 oz = __ldg(recs + 2 * (size_t)e.block), oo = __ldg(recs + 2 * (size_t)e.block + 1);
      const double own_zmin = __hiloint2double((int)oz.y, (int)oz.x), own_zmax = __hiloint2double((int)oz.w, (int)oz.z);
      const uint32_t own_model = oo.z, own_root = oo.w & kRecRootMask;
      const int32_t rix = (x >> kRBits) - g.rx0, riy = (y >> kRBits) - g.ry0;
      const uint32_t region = (uint32_t)(riy * g.nrx + rix);
      const uint32_t key = (region << (2 * kRBits)) | (((uint32_t)y & (kRegionWidth - 1)) << kRBits) | ((uint32_t)x & (kRegionWidth - 1));
      unsigned long long best_seq = ~0ull;
      uint32_t h = slot_home(key, g.slot_mask);
      for (uint32_t probes = 0; probes <= g.slot_mask; probes++) {
        const unsigned long long cur = __ldg(slots + h);
        if (cur == kSlotEmpty)
          break;
        if ((uint32_t)(cur >> 32) == key) {
          const uint32_t blk = (uint32_t)cur - 1u;
          const uint4 z = __ldg(recs + 2 * (size_t)blk), o = __ldg(recs + 2 * (size_t)blk + 1);
          const double zmin = __hiloint2double((int)z.y, (int)z.x), zmax = __hiloint2double((int)z.w, (int)z.z);
          // block.cc:153-157
          if (o.z != own_model && (o.w & kRecObstacle) && (o.w & kRecRootMask) != own_root && zmin <= own_zmax && zmax >= own_zmin) {
            const unsigned long long s = ((unsigned long long)o.y << 32) | o.x;
            if (s < best_seq) {
              best_seq = s;
              mine = (int32_t)o.z;
            }
          }
        }
        h = (h + 1) & g.slot_mask;
      }
    }
    const uint32_t votes = __ballot_sync(0xffffffffu, mine >= 0);
    if (votes)
      result = __shfl_sync(0xffffffffu, mine, __ffs(votes) - 1);
  }
  if (lane == 0)
    b.hit_model[warp] = result;
}

// ---- K4b: Model::AppendTouchingModels (model.cc:661-664 -> block.cc:116-127) -------------------------
// The set of models, not related to the query's own, that have a block in any cell the query's blocks are
// rendered into. The same cell walk as K4 (one warp per query, lane i looks at cell visit t0 + i), but no
// obstacle_return and no z test, and every unrelated entry counts, not the first. The warp keeps the set in
// shared memory: candidates are handed over one at a time (ballot order), the lanes search the list in
// parallel, lane 0 appends what is new. Written out ascending (the reference's std::set is ordered by
// pointer; what callers get from it is membership).
__global__ void __launch_bounds__(128) k4_touchers(GridView g, ToucherBatchView b)
{
  __shared__ uint32_t list_all[4][kTouchersCap];
  const uint32_t wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * 4 + wic;
  if (warp >= b.n_queries)
    return;
  uint32_t *list = list_all[wic];
  uint32_t count = 0; // warp-uniform
  bool overflow = false;
  const CollisionQuery q = b.queries[warp];
  const EdgeRec *edges = b.edges + q.first_edge;
  const int L = q.layer & 1;
  const unsigned long long *slots = g.slots[L];
  const uint4 *recs = reinterpret_cast<const uint4 *>(g.blk_rec[L]);
  for (uint32_t t0 = 0; t0 < q.n_steps; t0 += 32) {
    const uint32_t t = t0 + lane;
    uint32_t key = 0, own_root = 0, h = 0;
    bool live = t < q.n_steps;
    if (live) {
      uint32_t ei = 0;
      while (ei + 1 < q.n_edges && edges[ei + 1].first_step <= t)
        ei++;
      const EdgeRec e = edges[ei];
      int32_t x, y;
      edge_cell_post(e, t - e.first_step, x, y);
      own_root = __ldg(recs + 2 * (size_t)e.block + 1).w & kRecRootMask;
      const int32_t rix = (x >> kRBits) - g.rx0, riy = (y >> kRBits) - g.ry0;
      const uint32_t region = (uint32_t)(riy * g.nrx + rix);
      key = (region << (2 * kRBits)) | (((uint32_t)y & (kRegionWidth - 1)) << kRBits) | ((uint32_t)x & (kRegionWidth - 1));
      h = slot_home(key, g.slot_mask);
    }
    uint32_t probes = 0;
    for (;;) {
      // every lane advances along its cell's probe sequence to the next unrelated entry, if there is one
      uint32_t cand = 0xffffffffu;
      while (live) {
        const unsigned long long cur = __ldg(slots + h);
        if (cur == kSlotEmpty || probes > g.slot_mask) {
          live = false;
          break;
        }
        h = (h + 1) & g.slot_mask;
        probes++;
        if ((uint32_t)(cur >> 32) == key) {
          const uint4 o = __ldg(recs + 2 * (size_t)((uint32_t)cur - 1u) + 1);
          if ((o.w & kRecRootMask) != own_root) { // !IsRelated, block.cc:123
            cand = o.z;
            break;
          }
        }
      }
      uint32_t votes = __ballot_sync(0xffffffffu, cand != 0xffffffffu);
      if (!votes)
        break;
      while (votes) {
        const uint32_t m = __shfl_sync(0xffffffffu, cand, __ffs(votes) - 1);
        votes &= votes - 1u;
        bool seen = false;
        for (uint32_t i = lane; i < count; i += 32)
          seen |= list[i] == m;
        if (!__any_sync(0xffffffffu, seen)) {
          if (count < kTouchersCap) {
            if (lane == 0)
              list[count] = m;
            count++;
          } else {
            overflow = true;
          }
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  // ascending: the rank of an element is the number of smaller ones (the list holds no duplicates)
  uint32_t *out = b.out + (size_t)warp * b.max_per_query;
  for (uint32_t i = lane; i < count; i += 32) {
    const uint32_t m = list[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < count; j++)
      rank += list[j] < m;
    if (rank < b.max_per_query)
      out[rank] = m;
  }
  if (lane == 0)
    b.out_count[warp] = overflow ? 0xffffffffu : count;
}

// ---- K5: what the movers' new outlines touch (stg_rt_models_move) ----------------------------------
// The same cell walk and the same condition as K4, but instead of the first model in the way it
// reports (a) whether anything that does not move this tick is in the way and (b) which other movers
// are. The outline walked is the one the mover WANTS (it need not be rendered); the block's z range
// and ownership come from the edge records' block. One warp per mover.
__global__ void __launch_bounds__(128) k5_touches(GridView g, TouchBatchView b)
{
  __shared__ uint32_t list_all[4][kTouchMax];
  __shared__ uint32_t count_all[4], static_all[4];
  const uint32_t wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * 4 + wic;
  if (warp >= b.n_queries)
    return;
  uint32_t *list = list_all[wic];
  if (lane == 0)
    count_all[wic] = static_all[wic] = 0;
  __syncwarp();
  const CollisionQuery q = b.queries[warp];
  const EdgeRec *edges = b.edges + q.first_edge;
  const int L = q.layer & 1;
  const unsigned long long *slots = g.slots[L];
  const uint4 *recs = reinterpret_cast<const uint4 *>(g.blk_rec[L]);
  for (uint32_t t = lane; t < q.n_steps; t += 32) {
    uint32_t ei = 0;
    while (ei + 1 < q.n_edges && edges[ei + 1].first_step <= t)
      ei++;
    const EdgeRec e = edges[ei];
    int32_t x, y;
    edge_cell_post(e, t - e.first_step, x, y);
    const uint4 oo = __ldg(recs + 2 * (size_t)e.block + 1);
    const double own_zmin = b.edge_z[2 * (q.first_edge + ei)], own_zmax = b.edge_z[2 * (q.first_edge + ei) + 1];
    const uint32_t own_model = oo.z, own_root = oo.w & kRecRootMask;
    const int32_t rix = (x >> kRBits) - g.rx0, riy = (y >> kRBits) - g.ry0;
    const uint32_t region = (uint32_t)(riy * g.nrx + rix);
    const uint32_t key = (region << (2 * kRBits)) | (((uint32_t)y & (kRegionWidth - 1)) << kRBits) | ((uint32_t)x & (kRegionWidth - 1));
    uint32_t h = slot_home(key, g.slot_mask);
    for (uint32_t probes = 0; probes <= g.slot_mask; probes++) {
      const unsigned long long cur = __ldg(slots + h);
      if (cur == kSlotEmpty)
        break;
      if ((uint32_t)(cur >> 32) == key) {
        const uint32_t blk = (uint32_t)cur - 1u;
        const uint4 z = __ldg(recs + 2 * (size_t)blk), o = __ldg(recs + 2 * (size_t)blk + 1);
        const double zmin = __hiloint2double((int)z.y, (int)z.x), zmax = __hiloint2double((int)z.w, (int)z.z);
        if (o.z != own_model && (o.w & kRecObstacle) && (o.w & kRecRootMask) != own_root && zmin <= own_zmax && zmax >= own_zmin) {
          const uint32_t mover = __ldg(b.mover_of_model + o.z); // block.cc:153-157 holds: who is it?
          if (mover == 0) {
            static_all[wic] = 1;
          } else {
            bool seen = false;
            const uint32_t n = min(*(volatile uint32_t *)(count_all + wic), kTouchMax);
            for (uint32_t i = 0; i < n && !seen; i++)
              seen = ((volatile uint32_t *)list)[i] == mover - 1u;
            if (!seen) {
              const uint32_t at = atomicAdd(count_all + wic, 1u);
              if (at < kTouchMax)
                list[at] = mover - 1u;
            }
          }
        }
      }
      h = (h + 1) & g.slot_mask;
    }
  }
  __syncwarp();
  TouchRec *out = b.out + warp;
  if (lane == 0) {
    out->static_hit = static_all[wic];
    out->n_touched = count_all[wic];
  }
  if (lane < kTouchMax)
    out->touched[lane] = lane < min(count_all[wic], kTouchMax) ? list[lane] : 0u;
}

// ColorMatchIgnoreAlpha, model_blobfinder.cc:109-113
__device__ __forceinline__ bool color_match(const double *a, const double *c)
{
  const double epsilon = 1e-5;
  return fabs(a[0] - c[0]) < epsilon && fabs(a[1] - c[1]) < epsilon && fabs(a[2] - c[2]) < epsilon;
}

// One thread per blobfinder: the segmentation loop is sequential over the scan line (80 samples by
// default) and tiny next to the rays that produced it.
__global__ void __launch_bounds__(64) k3_blobs(GridView g, BlobBatchView b)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= b.n_finders)
    return;
  const BlobFinderRec bf = b.finders[f];
  const uint32_t first = b.fan_first_beam[f], w = bf.scan_width;
  const int32_t *mod = b.hit_model + first;
  const double *rng = b.ranges + first;
  const double yRadsPerPixel = bf.fov / bf.scan_height; // :173
  BlobRec *out = b.out + (size_t)f * b.max_per_finder;
  uint32_t n_out = 0;
  for (uint32_t s = 0; s < w; s++) { // :178
    if (mod[s] < 0)
      continue;
    const uint32_t right = s;
    const int32_t blob_model = mod[s];
    const double *blobcol = b.mdl_color + 4 * (size_t)blob_model;
    while (s < w && mod[s] >= 0 && color_match(b.mdl_color + 4 * (size_t)mod[s], blobcol)) // :190
      s++;
    const uint32_t left = s - 1;
    if (bf.n_colors) { // :201-212
      bool found = false;
      for (uint32_t c = 0; c < bf.n_colors; c++)
        if (color_match(blobcol, b.colors + 3 * (size_t)(bf.first_color + c))) {
          found = true;
          break;
        }
      if (!found)
        continue;
    }
    const double robotHeight = 0.6;
    double range = 0;
    for (uint32_t t = right; t <= left; t++)
      range += rng[t];
    range /= left - right + 1; // :221 (unsigned count converted to double)
    const double startyangle = atan2(robotHeight / 2.0, range);
    const double endyangle = -startyangle;
    int blobtop = (int)(bf.scan_height / 2) - (int)(startyangle / yRadsPerPixel);
    int blobbottom = (int)(bf.scan_height / 2) - (int)(endyangle / yRadsPerPixel);
    blobtop = max(blobtop, 0);
    blobbottom = min(blobbottom, (int)bf.scan_height);
    if (n_out < b.max_per_finder) {
      BlobRec r;
      r.model = (uint32_t)blob_model;
      r.left = w - left - 1;
      r.top = (uint32_t)blobtop;
      r.right = w - right - 1;
      r.bottom = (uint32_t)blobbottom;
      r.pad = 0;
      r.range = range;
      out[n_out] = r;
    }
    n_out++;
  }
  b.out_count[f] = n_out;
}

// diagnostics: the device's cos / sin of arbitrary headings (tests/test_gpu_trig.py)
__global__ void __launch_bounds__(256) k_debug_trig(uint64_t n, const double *x, double *cs)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    trace::heading_trig(x[i], cs[2 * i], cs[2 * i + 1]);
}

} // namespace

void launch_debug_trig(uint64_t n, const double *x, double *cs, void *stream)
{
  if (n)
    k_debug_trig<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, cs);
}

void launch_fiducials(const GridView &g, const FidBatchView &b, void *stream)
{
  if (!b.n_finders)
    return;
  const unsigned blocks = (b.n_finders * 32 + 127) / 128;
  k3_fiducials<<<blocks, 128, 0, (cudaStream_t)stream>>>(g, b);
}

// v.start must be zeroed (n_bins + 1 words) before the call
void launch_fiducials_grid(const GridView &g, const FidBatchView &b, const FidGridView &v, void *stream)
{
  if (!b.n_finders)
    return;
  cudaStream_t s = (cudaStream_t)stream;
  if (b.n_targets) {
    k3_fid_bin_count<<<(b.n_targets + 255) / 256, 256, 0, s>>>(v, b.targets, b.n_targets);
    k3_fid_bin_scan<<<1, 1024, 0, s>>>(v);
    k3_fid_bin_scatter<<<(b.n_targets + 255) / 256, 256, 0, s>>>(v, b.n_targets);
  } else {
    k3_fid_bin_scan<<<1, 1024, 0, s>>>(v);
  }
  k3_fiducials_grid<<<(b.n_finders * 32 + 127) / 128, 128, 0, s>>>(g, b, v);
}

// the line-of-sight ray of AddModelIfVisible (model_fiducial.cc:179) as a fan record per finder, derived from the finder
// records where they lie (100,000 finders: 1 ms of host time and 6.4 MB of upload otherwise)
__global__ void __launch_bounds__(256) k3_fid_fans(const FidFinderRec *finders, FanRec *fans, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const FidFinderRec &f = finders[i];
  FanRec r;
  r.finder = f.finder;
  r.n_samples = 1;
  r.pred = 1; // STG_RT_PRED_UNRELATED (fiducial_raytrace_match, model_fiducial.cc:103-107)
  r.ztest = 1;
  r.layer = f.layer;
  r.angles = 2; // STG_RT_ANGLES_EXPLICIT
  r.first_heading = 0;
  r.ox = f.ox;
  r.oy = f.oy;
  r.oz = f.oz;
  r.oa = f.oa;
  r.range = f.max_range_anon;
  r.fov = 0.0;
  fans[i] = r;
}

void launch_fid_fans(const FidFinderRec *finders, FanRec *fans, uint32_t n, void *stream)
{
  if (n)
    k3_fid_fans<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(finders, fans, n);
}

void launch_fid_pack(const FidBatchView &b, uint32_t *offsets, void *packed_host, uint32_t *count_host, void *stream)
{
  if (!b.n_finders)
    return;
  cudaStream_t s = (cudaStream_t)stream;
  k3_fid_pack_scan<<<1, 1024, 0, s>>>(b, offsets);
  k3_fid_pack<<<(b.n_finders * 32 + 127) / 128, 128, 0, s>>>(b, offsets, (unsigned long long *)packed_host, count_host);
}

void launch_collisions(const GridView &g, const CollisionBatchView &b, void *stream)
{
  if (!b.n_queries)
    return;
  k4_collisions<<<(b.n_queries * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(g, b);
}

void launch_touchers(const GridView &g, const ToucherBatchView &b, void *stream)
{
  if (!b.n_queries)
    return;
  k4_touchers<<<(b.n_queries + 3) / 4, 128, 0, (cudaStream_t)stream>>>(g, b);
}

// table[pairs[2i]] = clear ? 0 : pairs[2i + 1]
__global__ void __launch_bounds__(256) k5_mark_movers(const uint32_t *pairs, uint32_t n, uint32_t *table, int clear)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    table[pairs[2 * i]] = clear ? 0u : pairs[2 * i + 1];
}

void launch_mark_movers(const uint32_t *pairs, uint32_t n, uint32_t *table, int clear, void *stream)
{
  if (n)
    k5_mark_movers<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pairs, n, table, clear);
}

void launch_touches(const GridView &g, const TouchBatchView &b, void *stream)
{
  if (!b.n_queries)
    return;
  k5_touches<<<(b.n_queries + 3) / 4, 128, 0, (cudaStream_t)stream>>>(g, b);
}

void launch_blobs(const GridView &g, const BlobBatchView &b, void *stream)
{
  if (!b.n_finders)
    return;
  k3_blobs<<<(b.n_finders + 63) / 64, 64, 0, (cudaStream_t)stream>>>(g, b);
}

} // namespace stg
