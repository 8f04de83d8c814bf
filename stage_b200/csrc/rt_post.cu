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
    bool accept = false;
    FiducialRec rec;
    if (t < b.n_targets) {
      const FidTargetRec him = b.targets[t];
      // the x / y window of the sorted-set query (model_fiducial.cc:245-262), bounds inclusive
      bool cand = him.model != fd.finder && him.x >= fd.x - fd.max_range_anon && him.x <= fd.x + fd.max_range_anon &&
                  him.y >= fd.y - fd.max_range_anon && him.y <= fd.y + fd.max_range_anon;
      cand = cand && him.fiducial_key == fd.fiducial_key; // :117
      double range = 0, dtheta = 0;
      if (cand) {
        const double dx = him.x - fd.x, dy = him.y - fd.y;
        range = hypot(dy, dx);                                   // :130
        cand = !(range >= fd.max_range_anon);                    // :134
        dtheta = normalize_angle(atan2(dy, dx) - fd.a);          // :144-145
        cand = cand && !(fabs(dtheta) > fd.fov / 2.0);           // :147
        cand = cand && __ldg(g.mdl_root + him.model) != finder_root; // IsRelated, :152
      }
      if (cand) {
        // Raytrace(Pose(0,0,0,dtheta), max_range_anon, fiducial_raytrace_match, NULL, true), :179:
        // LocalToGlobal adds the heading onto (GetGlobalPose() + geom.pose) and normalises it
        double cosa, sina;
        trace::heading_trig(normalize_angle(fd.oa + dtheta), cosa, sina);
        trace::RayResult r;
        trace::trace_ray<false>(g, fan, cosa, sina, r);
        int32_t seen = r.hit_model;
        if (fd.ignore_zloc && seen < 0) // :184
          seen = (int32_t)him.model;
        if (seen == (int32_t)him.model) { // :192
          accept = true;
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
      }
    }
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

void launch_blobs(const GridView &g, const BlobBatchView &b, void *stream)
{
  if (!b.n_finders)
    return;
  k3_blobs<<<(b.n_finders + 63) / 64, 64, 0, (cudaStream_t)stream>>>(g, b);
}

} // namespace stg
