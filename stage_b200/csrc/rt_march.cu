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
//    test per cell. Only where a mask bit is set is the cell's content looked up.
//  * Empty regions are skipped with the reference's own crossing arithmetic (FP64, same operation
//    order, no FMA contraction: build with -fmad=false), including its stale error term, the
//    truncating double->int conversions and the int -= double of the remaining distance.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_device.h"

namespace stg {

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t k)
{
  k ^= k >> 16;
  k *= 0x7feb352du;
  k ^= k >> 15;
  k *= 0x846ca68bu;
  k ^= k >> 16;
  return k;
}

struct HitTest {
  uint32_t finder, finder_root;
  uint32_t pred, ztest;
  double oz;
};

// Resolve an occupied cell: among the blocks rendered into it that pass the z test and the
// predicate, the one first in Cell::blocks[layer] order wins (world.cc:841-862).
__device__ __noinline__ int32_t resolve_cell(const GridView &g, const HitTest &ht, int layer, uint32_t key)
{
  const unsigned long long *slots = g.slots[layer];
  const unsigned long long *seq = g.blk_seq[layer];
  unsigned long long best_seq = ~0ull;
  int32_t best_model = -1;
  uint32_t h = mix32(key) & g.slot_mask;
  for (uint32_t probes = 0; probes <= g.slot_mask; probes++) {
    const unsigned long long cur = __ldg(slots + h);
    if (cur == kSlotEmpty)
      break;
    if ((uint32_t)(cur >> 32) == key) {
      const uint32_t blk = (uint32_t)cur - 1u;
      const bool z_ok = !ht.ztest || !(ht.oz < __ldg(g.blk_zmin + blk) || ht.oz > __ldg(g.blk_zmax + blk));
      if (z_ok) {
        const uint32_t m = __ldg(g.blk_model + blk);
        bool pass;
        if (ht.pred == 0u) // ranger_match, model_ranger.cc:158-170
          pass = __ldg(g.mdl_root + m) != ht.finder_root && !(__ldg(g.mdl_ranger_return + m) < 0.0);
        else if (ht.pred == 2u) // gripper_raytrace_match, model_gripper.cc:329-336
          pass = m != ht.finder && (__ldg(g.mdl_flags + m) & 1u);
        else // fiducial / blob / bumper: not related, model.cc:446-466
          pass = __ldg(g.mdl_root + m) != ht.finder_root;
        if (pass) {
          const unsigned long long s = __ldg(seq + blk);
          if (s < best_seq) {
            best_seq = s;
            best_model = (int32_t)m;
          }
        }
      }
    }
    h = (h + 1) & g.slot_mask;
  }
  return best_model;
}

// The value the reference's `x += s` (s = +-1.0) reaches after k repetitions. Every one of those
// additions is exact as long as the running value stays within the binade of x (or below it): all
// intermediate values are multiples of ulp(x) and fit 53 bits. Then one addition of k*s gives the
// same double. Otherwise (the magnitude grows across a power of two, or the sign changes) the
// roundings happen one by one, as there.
__device__ __forceinline__ double seq_add(double x, int32_t k, int32_t s)
{
  const double r = x + (double)(k * s);
  const int ex = (__double2hiint(x) >> 20) & 0x7ff, er = (__double2hiint(r) >> 20) & 0x7ff;
  const bool same_sign = (__double2hiint(x) ^ __double2hiint(r)) >= 0;
  if ((same_sign && er <= ex) || k == 0)
    return r;
  const double ds = (double)s;
  for (int32_t j = 0; j < k; j++)
    x += ds;
  return x;
}

} // namespace

__global__ void __launch_bounds__(128, 6) k2_ray_march(GridView g, FanBatchView b)
{
  const uint64_t beam = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (beam >= b.n_beams)
    return;

  // which fan does this beam belong to
  uint32_t lo = 0, hi = b.n_fans;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(b.fan_first_beam + mid) <= beam)
      lo = mid;
    else
      hi = mid;
  }
  const FanRec fan = b.fans[lo];

  HitTest ht;
  ht.finder = fan.finder;
  ht.finder_root = __ldg(g.mdl_root + fan.finder);
  ht.pred = fan.pred;
  ht.ztest = fan.ztest;
  ht.oz = fan.oz;
  const int layer = fan.layer & 1;
  const double ppm = g.ppm;

  double gx = fan.ox * ppm, gy = fan.oy * ppm; // world.cc:765-766
  const double x_start = gx, y_start = gy;
  const double heading = b.heading[beam];
  const double ang = heading == 0.0 ? 1e-12 : heading; // world.cc:773
  double sina, cosa;
  if (b.trig_in) {
    cosa = b.trig_in[2 * beam];
    sina = b.trig_in[2 * beam + 1];
  } else {
    sina = sin(ang);
    cosa = cos(ang);
  }
  const double tana = sina / cosa;
  const double dx = ppm * fan.range * cosa, dy = ppm * fan.range * sina;
  const int32_t sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
  const int32_t ax = __double2int_rz(fabs(dx)), ay = __double2int_rz(fabs(dy));
  int32_t n = ax + ay; // manhattan distance left, world.cc:792

  // The reference's error term exy (ay-ax; exy<0: x step, exy += 2ay; else y step, exy -= 2ax) in
  // run form. A ray is walked along its major axis ("run" axis, coordinate a) in runs of cells that
  // share the cross coordinate b:  E < 0 -> run step, E += b_run;  E >= 0 -> cross step, E -= b_cross.
  //   x-major: E = exy,      b_run = 2ay, b_cross = 2ax, masks = row masks,    a = cx, b = cy
  //   y-major: E = -exy - 1, b_run = 2ax, b_cross = 2ay, masks = column masks, a = cy, b = cx
  // Both forms describe the same cell sequence for any slope; the choice only makes runs long.
  const bool ymajor = ay > ax;
  const int32_t b_run = ymajor ? 2 * ax : 2 * ay, b_cross = ymajor ? 2 * ay : 2 * ax;
  int32_t E = ymajor ? -(ay - ax) - 1 : (ay - ax);
  const int32_t s_run = ymajor ? sy : sx, s_cross = ymajor ? sx : sy;
  const float inv_run = b_run ? 1.0f / (float)b_run : 0.0f;
  const uint32_t mask_ofs = ymajor ? kRegionWidth : 0;

  // world.cc:795-802
  const double xjumpx = sx * kRegionWidth, xjumpy = sx * kRegionWidth * tana;
  const double yjumpx = sy * kRegionWidth / tana, yjumpy = sy * kRegionWidth;
  const double xjumpdist = fabs(xjumpx) + fabs(xjumpy);
  const double yjumpdist = fabs(yjumpx) + fabs(yjumpy);

  double xcrossx = 0, xcrossy = 0, ycrossx = 0, ycrossy = 0, distX = 0, distY = 0;
  bool need_crossings = true;

  double range = fan.range;
  int32_t hit_model = -1, hit_x = 0, hit_y = 0;
  uint32_t cell_visits = 0, region_probes = 0;
  const uint32_t *occ = g.occ[layer];

  while (n > 0) {
    region_probes++;
    const int32_t ix0 = __double2int_rz(gx), iy0 = __double2int_rz(gy); // GETSREG/GETREG(double), world.cc:820-821
    const int32_t grx = ix0 >> kRBits, gry = iy0 >> kRBits;
    const int32_t rix = grx - g.rx0, riy = gry - g.ry0;
    uint32_t region = 0, count = 0;
    if ((uint32_t)rix < (uint32_t)g.nrx && (uint32_t)riy < (uint32_t)g.nry) {
      region = (uint32_t)(riy * g.nrx + rix);
      count = __ldg(g.reg_count + region);
    }
    if (count) {
      // ---- populated region: integer walk over the occupancy masks (world.cc:823-882) ----
      need_crossings = true;
      const int32_t cx0 = ix0 & (kRegionWidth - 1), cy0 = iy0 & (kRegionWidth - 1);
      int32_t a = ymajor ? cy0 : cx0, c = ymajor ? cx0 : cy0; // run / cross coordinate in the region
      int32_t ka = 0, kc = 0;                                   // steps taken along each
      const uint32_t *tile = occ + (size_t)region * kTileWords + mask_ofs;
      uint32_t word = __ldg(tile + c);
      bool hit = false;
      for (;;) {
        // r = run steps before the next cross step: smallest r >= 0 with E + r*b_run >= 0
        int32_t r = 0;
        if (E < 0) {
          const int32_t t = -E;
          const float q = (float)t * inv_run;
          r = (b_run == 0 || q >= 33.0f) ? 33 : (int32_t)q;
          if (r < 33) {
            while (r * b_run < t)
              r++;
            while (r > 0 && (r - 1) * b_run >= t)
              r--;
          }
        }
        const int32_t room = s_run > 0 ? kRegionWidth - a : a + 1; // cells up to the region edge, this one included
        int32_t m = r + 1 < room ? r + 1 : room;
        m = m < n ? m : n; // a cell is only tested while n > 0 (world.cc:840)
        const uint32_t span = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
        const uint32_t msk = s_run > 0 ? span << a : span << (a - m + 1);
        const uint32_t occupied = word & msk;
        if (occupied) {
          const int32_t ah = s_run > 0 ? __ffs(occupied) - 1 : 31 - __clz(occupied);
          const int32_t j = s_run > 0 ? ah - a : a - ah;
          const uint32_t hx = ymajor ? (uint32_t)c : (uint32_t)ah, hy = ymajor ? (uint32_t)ah : (uint32_t)c;
          const int32_t mdl = resolve_cell(g, ht, layer, (region << (2 * kRBits)) | (hy << kRBits) | hx);
          if (mdl >= 0) {
            hit_model = mdl;
            hit_x = grx * kRegionWidth + (int32_t)hx;
            hit_y = gry * kRegionWidth + (int32_t)hy;
            ka += j;
            cell_visits += (uint32_t)j + 1u;
            hit = true;
            break;
          }
          word &= ~(1u << ah); // nothing in that cell stops this ray: look past it
          continue;
        }
        cell_visits += (uint32_t)m;
        n -= m;
        if (m == r + 1) { // the whole run, then the cross step
          a += s_run * r;
          ka += r;
          c += s_cross;
          kc += 1;
          E += r * b_run - b_cross;
          if ((uint32_t)c >= (uint32_t)kRegionWidth || n <= 0)
            break;
          word = __ldg(tile + c);
        } else { // cut short by the region edge or by the end of the ray
          ka += m;
          E += m * b_run;
          break;
        }
      }
      // bring the double position up to date (world.cc:868-877 did it step by step)
      gx = seq_add(gx, ymajor ? kc : ka, sx);
      gy = seq_add(gy, ymajor ? ka : kc, sy);
      if (hit) {
        if (ax > ay) // world.cc:856-859
          range = fabs((gx - x_start) / cosa) / ppm;
        else
          range = fabs((gy - y_start) / sina) / ppm;
        break;
      }
    } else {
      // ---- empty region: jump to the next region boundary (world.cc:884-957) ----
      if (need_crossings) {
        need_crossings = false;
        double regionx = (double)(ix0 / kRegionWidth * kRegionWidth);
        double regiony = (double)(iy0 / kRegionWidth * kRegionWidth);
        if ((gx < 0) && (ix0 % kRegionWidth))
          regionx -= kRegionWidth;
        if ((gy < 0) && (iy0 % kRegionWidth))
          regiony -= kRegionWidth;
        const double xdx = sx < 0 ? regionx - gx - 1.0 : regionx + kRegionWidth - gx;
        const double xdy = xdx * tana;
        const double ydy = sy < 0 ? regiony - gy - 1.0 : regiony + kRegionWidth - gy;
        const double ydx = ydy / tana;
        xcrossx = gx + xdx;
        xcrossy = gy + xdy;
        ycrossx = gx + ydx;
        ycrossy = gy + ydy;
        distX = fabs(xdx) + fabs(xdy);
        distY = fabs(ydx) + fabs(ydy);
      }
      if (distX < distY) {
        gx = xcrossx;
        gy = xcrossy;
        n = __double2int_rz((double)n - distX); // int -= double, world.cc:931
        xcrossx += xjumpx;
        xcrossy += xjumpy;
        distY -= distX;
        distX = xjumpdist;
      } else {
        gx = ycrossx;
        gy = ycrossy;
        n = __double2int_rz((double)n - distY);
        ycrossx += yjumpx;
        ycrossy += yjumpy;
        distX -= distY;
        distY = yjumpdist;
      }
    }
  }

  // K3 (ranger epilogue, fused): model_ranger.cc:243-251 with zero noise
  if (b.ranges)
    b.ranges[beam] = range;
  if (b.intensities)
    b.intensities[beam] = hit_model >= 0 ? __ldg(g.mdl_ranger_return + hit_model) : 0.0;
  if (b.hit_model)
    b.hit_model[beam] = hit_model;
  if (b.hit_cell) {
    b.hit_cell[2 * beam] = hit_x;
    b.hit_cell[2 * beam + 1] = hit_y;
  }
  if (b.steps) {
    b.steps[2 * beam] = cell_visits;
    b.steps[2 * beam + 1] = region_probes;
  }
}

void launch_ray_march(const GridView &g, const FanBatchView &b, void *stream)
{
  if (!b.n_beams)
    return;
  const uint64_t blocks = (b.n_beams + 127) / 128;
  k2_ray_march<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(g, b);
}

} // namespace stg
