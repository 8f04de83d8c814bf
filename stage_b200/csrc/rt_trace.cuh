// Device-side ray tracer shared by the march kernel (rt_march.cu) and the sensor post-process kernels
// (rt_post.cu). See rt_march.cu for the design notes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_device.h"
#include "rt_libm.cuh"
#include "rt_walk.h"

namespace stg {
namespace trace {

// Resolve an occupied cell: among the blocks rendered into it that pass the z test and the
// predicate, the one first in Cell::blocks[layer] order wins (world.cc:841-862). Reads the fan's
// constants from its (L1-resident) record rather than holding them in registers during the march.
// The probe sequence is read a 32-byte sector (4 slots) at a time; it ends at the first empty slot.
static __device__ __forceinline__ int32_t resolve_cell(const unsigned long long *slots, uint32_t slot_mask, const uint4 *recs,
                                             const uint32_t *mdl_root, const FanRec *fan, uint32_t key)
{
  const uint32_t finder = fan->finder, pred = fan->pred, ztest = fan->ztest;
  const double oz = fan->oz;
  const uint32_t finder_root = __ldg(mdl_root + finder);
  unsigned long long best_seq = ~0ull;
  int32_t best_model = -1;
  uint32_t h = slot_home(key, slot_mask);
  for (uint32_t probes = 0; probes <= slot_mask; probes += 4) {
    const ulonglong2 p01 = __ldg(reinterpret_cast<const ulonglong2 *>(slots + h));
    const ulonglong2 p23 = __ldg(reinterpret_cast<const ulonglong2 *>(slots + h + 2));
    // which of the four precede the first empty slot and hold this cell
    const bool e0 = p01.x == kSlotEmpty, e1 = e0 || p01.y == kSlotEmpty, e2 = e1 || p23.x == kSlotEmpty,
               e3 = e2 || p23.y == kSlotEmpty;
    uint32_t match = (!e0 && (uint32_t)(p01.x >> 32) == key ? 1u : 0u) | (!e1 && (uint32_t)(p01.y >> 32) == key ? 2u : 0u) |
                     (!e2 && (uint32_t)(p23.x >> 32) == key ? 4u : 0u) | (!e3 && (uint32_t)(p23.y >> 32) == key ? 8u : 0u);
    while (match) {
      const uint32_t j = __ffs(match) - 1;
      match &= match - 1u;
      const unsigned long long cur = j == 0 ? p01.x : j == 1 ? p01.y : j == 2 ? p23.x : p23.y;
      const uint32_t blk = (uint32_t)cur - 1u;
      const uint4 z = __ldg(recs + 2 * (size_t)blk), o = __ldg(recs + 2 * (size_t)blk + 1);
      const double zmin = __hiloint2double((int)z.y, (int)z.x), zmax = __hiloint2double((int)z.w, (int)z.z);
      if (!ztest || !(oz < zmin || oz > zmax)) { // world.cc:846
        const uint32_t model = o.z, root = o.w & kRecRootMask;
        bool pass;
        if (pred == 0u) // ranger_match, model_ranger.cc:158-170
          pass = root != finder_root && !(o.w & kRecNoRanger);
        else if (pred == 2u) // gripper_raytrace_match, model_gripper.cc:329-336
          pass = model != finder && (o.w & kRecGripper);
        else // fiducial / blob / bumper: not related, model.cc:446-466
          pass = root != finder_root;
        if (pass) {
          const unsigned long long s = ((unsigned long long)o.y << 32) | o.x;
          if (s < best_seq) {
            best_seq = s;
            best_model = (int32_t)model;
          }
        }
      }
    }
    if (e3)
      break;
    h = (h + 4) & slot_mask;
  }
  return best_model;
}

// The value the reference's `x += s` (s = +-1.0) reaches after k repetitions. Every one of those
// additions is exact as long as the running value stays within the binade of x (or below it): all
// intermediate values are multiples of ulp(x) and fit 53 bits. Then one addition of k*s gives the
// same double. Otherwise (the magnitude grows across a power of two, or the sign changes) the
// roundings happen one by one, as there.
__device__ __forceinline__ double seq_add(double x, int32_t k, bool negative)
{
  const double ks = negative ? -(double)k : (double)k;
  const double r = x + ks;
  const int hx = __double2hiint(x), hr = __double2hiint(r);
  if (k == 0 || ((hx ^ hr) >= 0 && (hr & 0x7ff00000) <= (hx & 0x7ff00000)))
    return r;
  const double ds = negative ? -1.0 : 1.0;
  for (int32_t j = 0; j < k; j++)
    x += ds;
  return x;
}

// m bits from bit a on (0 <= a, 1 <= m, a + m <= 32): one BMSK instead of two shifts and their operands
__device__ __forceinline__ uint32_t run_mask(int32_t a, int32_t m)
{
  uint32_t mask;
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(mask) : "r"(a), "r"(m));
  return mask;
}

// Which region holds the position (gx, gy), and Region::count of it (0: no such region, or nothing in it):
// the reference truncates the double position toward zero (GETSREG / GETREG(double), world.cc:820-821), so cells
// at -1 < x < 0 alias cell 0.
__device__ __forceinline__ uint32_t probe_region(const GridView &g, double gx, double gy, int32_t &ix0, int32_t &iy0, uint32_t &region)
{
  ix0 = __double2int_rz(gx);
  iy0 = __double2int_rz(gy);
  const int32_t rix = (ix0 >> kRBits) - g.rx0, riy = (iy0 >> kRBits) - g.ry0;
  region = 0;
  if ((uint32_t)rix < (uint32_t)g.nrx && (uint32_t)riy < (uint32_t)g.nry) {
    region = (uint32_t)(riy * g.nrx + rix);
    return __ldg(&g.head[region].count);
  }
  return 0;
}

// What the walk of one populated region needs besides its running state.
struct WalkConst {
  int32_t b_run, b_cross, rm1, v1, cp_ahead;
  uint32_t revmask, aflip, cflip;
  bool ymajor;
};

// The walk of one populated region from (ap, cp) with error term E, r run steps to go before the next cross
// step, `word` the direction-normalised mask of the current row / column; `tile` = word index of the region's
// row (column) masks in GridView::occ, a multiple of 32: mask number (cp ^ cflip) belongs to cross coordinate cp.
// Walk and look-up alternate: the walk tests run after run until one holds an occupied cell and leaves its
// loop there; the lanes of the warp that met something in this region then look their cells up side by side
// (one wait for the cell-entry sectors instead of one per lane), and a lane whose cell does not stop it goes
// back to walking. Returns the model that stops the ray (ap, cp = its cell) or -1 when the region or the ray
// has ended (ap, cp, E as the reference's cell-by-cell loop leaves them). kBounded: the ray may end inside
// this region, n is kept up to date; otherwise n is not touched.
template <bool kSteps, bool kBounded>
__device__ __forceinline__ int32_t walk_region(const GridView &g, const FanRec *fan, bool layer1, uint32_t region, const WalkConst &wc,
                                               int32_t &ap, int32_t &cp, int32_t &E, int32_t &n, int32_t r, uint32_t word,
                                               uint32_t tile, uint32_t &cell_visits)
{
  uint32_t occupied = 0, ahead = 0;
  int32_t m = 0;
  bool more = true; // false: the region, or the ray, has ended
  for (;;) {
    for (;;) {
      // the mask the walk needs after this run's cross step, requested before this run is tested
      ahead = 0;
      if (cp < wc.cp_ahead)
        ahead = __ldg(g.occ[0] + (tile | ((uint32_t)(cp + 1) ^ wc.cflip))); // one LOP3, one wide multiply-add
      m = min(r + 1, kRegionWidth - ap);
      if (kBounded)
        m = min(m, n); // cells tested: a cell is only tested while n > 0
      occupied = word & run_mask(ap, m);
      if (occupied)
        break;
      if (kSteps)
        cell_visits += (uint32_t)m;
      if (kBounded)
        n -= m;
      if (m != r + 1) { // cut short by the region edge or by the end of the ray
        ap += m;
        E += m * wc.b_run;
        more = false;
        break;
      }
      // the whole run, then the cross step
      ap += r;
      cp += 1;
      E += r * wc.b_run - wc.b_cross;
      if (cp >= kRegionWidth || (kBounded && n <= 0)) {
        more = false;
        break;
      }
      word = ahead ^ ((ahead ^ __brev(ahead)) & wc.revmask);
      r = wc.rm1 + (int32_t)((uint32_t)(wc.v1 + E) >> 31); // v1 >= -E ? r_max - 1 : r_max
    }
    if (!more)
      return -1;
    do {
      const int32_t ah = __ffs(occupied) - 1;
      const uint32_t areal = (uint32_t)ah ^ wc.aflip, creal = (uint32_t)cp ^ wc.cflip;
      const uint32_t hx = wc.ymajor ? creal : areal, hy = wc.ymajor ? areal : creal;
      const int32_t mdl = resolve_cell(layer1 ? g.slots[1] : g.slots[0], g.slot_mask,
                                       reinterpret_cast<const uint4 *>(layer1 ? g.blk_rec[1] : g.blk_rec[0]),
                                       g.mdl_root, fan, (region << (2 * kRBits)) | (hy << kRBits) | hx);
      if (mdl >= 0) {
        if (kSteps)
          cell_visits += (uint32_t)(ah - ap) + 1u;
        ap = ah;
        return mdl;
      }
      occupied &= occupied - 1u; // nothing in that cell stops this ray: look past it
    } while (occupied);
    // nothing in this run stops the ray: past it, as above
    if (kSteps)
      cell_visits += (uint32_t)m;
    if (kBounded)
      n -= m;
    if (m != r + 1) {
      ap += m;
      E += m * wc.b_run;
      return -1;
    }
    ap += r;
    cp += 1;
    E += r * wc.b_run - wc.b_cross;
    if (cp >= kRegionWidth || (kBounded && n <= 0))
      return -1;
    word = ahead ^ ((ahead ^ __brev(ahead)) & wc.revmask);
    r = wc.rm1 + (int32_t)((uint32_t)(wc.v1 + E) >> 31);
  }
}

enum : uint32_t { kYMajor = 1u, kRunNeg = 2u, kCrossNeg = 4u, kXNeg = 8u, kYNeg = 16u, kClosedForms = 64u };

struct RayResult {
  double range;
  int32_t hit_model, hit_x, hit_y;
  uint32_t cell_visits, region_probes;
};

// One World::Raytrace(Ray) (world.cc:756-963). fan supplies finder / predicate / z test / layer /
// origin / range; cosa, sina are cos and sin of the heading the reference would have used
// (world.cc:773-775, after its zero-heading substitution).
template <bool kSteps>
__device__ __forceinline__ void trace_ray(const GridView &g, const FanRec *fan, double cosa, double sina, RayResult &res)
{
  const double ppm = g.ppm;
  double gx = fan->ox * ppm, gy = fan->oy * ppm; // world.cc:765-766
  double tana, trig_major, yjumpx;
  int32_t n, E, b_run, b_cross, r_max, v1;
  uint32_t flags;
  {
    tana = sina / cosa;
    const double range = fan->range;
    const double dx = ppm * range * cosa, dy = ppm * range * sina; // world.cc:780-781
    const int32_t ax = __double2int_rz(fabs(dx)), ay = __double2int_rz(fabs(dy));
    n = ax + ay; // manhattan distance left, world.cc:792
    // The reference's error term exy (ay-ax; exy<0: x step, exy += 2ay; else y step, exy -= 2ax) in
    // run form. A ray is walked along its major ("run") axis in runs of cells that share the cross
    // coordinate:  E < 0 -> run step, E += b_run;  E >= 0 -> cross step, E -= b_cross.
    //   x-major: E = exy,      b_run = 2ay, b_cross = 2ax, row masks
    //   y-major: E = -exy - 1, b_run = 2ax, b_cross = 2ay, column masks
    // Both forms describe the same cell sequence for any slope; the choice only makes runs long.
    const bool ymajor = ay > ax;
    b_run = ymajor ? 2 * ax : 2 * ay;
    b_cross = ymajor ? 2 * ay : 2 * ax;
    E = ymajor ? -(ay - ax) - 1 : (ay - ax);
    const bool xneg = dx < 0, yneg = dy < 0;
    flags = (ymajor ? kYMajor : 0u) | ((ymajor ? yneg : xneg) ? kRunNeg : 0u) | ((ymajor ? xneg : yneg) ? kCrossNeg : 0u) |
            (xneg ? kXNeg : 0u) | (yneg ? kYNeg : 0u) | (b_cross <= 2 * walk::kMaxMajorCells ? kClosedForms : 0u);
    // r_max = min(ceil(b_cross / b_run), 33): the longest possible run; 33 = "past the region edge"
    r_max = 33;
    if (b_run > 0 && (long long)b_run * 33 > b_cross)
      r_max = (b_cross + b_run - 1) / b_run;
    v1 = (r_max - 1) * b_run;
    trig_major = ax > ay ? cosa : sina; // the divisor of the hit range, world.cc:856-859
    yjumpx = (yneg ? -(double)kRegionWidth : (double)kRegionWidth) / tana; // world.cc:797
  }

  int32_t hit_model = -1, hit_x = 0, hit_y = 0;
  uint32_t cell_visits = 0, region_probes = 0;
  const bool layer1 = (fan->layer & 1) != 0;
  // the direction bits as the loops want them (taken from the packed word, not from dx / dy, so that
  // those doubles are dead after the set-up)
  asm volatile("" : "+r"(flags));
  const bool ymajor = (flags & kYMajor) != 0, runneg = (flags & kRunNeg) != 0;
  const bool xneg = (flags & kXNeg) != 0, yneg = (flags & kYNeg) != 0;
  const uint32_t aflip = runneg ? kRegionWidth - 1 : 0u, cflip = (flags & kCrossNeg) ? kRegionWidth - 1 : 0u;
  const uint32_t revmask = runneg ? ~0u : 0u;
  const int32_t rm1 = r_max - 1;
  const uint32_t occ_layer = layer1 ? (uint32_t)(g.occ[1] - g.occ[0]) : 0u; // one allocation, rt_api.cu
  // The tag of the regions this ray cannot be stopped in (GridView::reg_owner): regions that hold nothing but
  // blocks of the finder's own tree, under a predicate that rejects related models.
  const uint32_t finder_tag = fan->pred == 2u ? kOwnerNever : __ldg(g.mdl_root + fan->finder) + 1u;
  bool hit = false;

  while (n > 0) {
    if (kSteps)
      region_probes++;
    int32_t ix0, iy0;
    uint32_t region;
    uint32_t count = probe_region(g, gx, gy, ix0, iy0, region);
    if (!count) {
      // ---- a stretch of empty regions: jump from region boundary to region boundary (world.cc:884-957) until a
      // populated region or the end of the ray. The crossing state is set up at the first of them (the
      // reference's calculatecrossings, true after every populated region) and lives only here.
      double xcrossx, xcrossy, ycrossx, ycrossy, distX, distY;
      {
        double regionx = (double)(ix0 / kRegionWidth * kRegionWidth);
        double regiony = (double)(iy0 / kRegionWidth * kRegionWidth);
        if ((gx < 0) && (ix0 % kRegionWidth))
          regionx -= kRegionWidth;
        if ((gy < 0) && (iy0 % kRegionWidth))
          regiony -= kRegionWidth;
        const double xdx = (flags & kXNeg) ? regionx - gx - 1.0 : regionx + kRegionWidth - gx;
        const double xdy = xdx * tana;
        const double ydy = (flags & kYNeg) ? regiony - gy - 1.0 : regiony + kRegionWidth - gy;
        const double ydx = ydy / tana;
        xcrossx = gx + xdx;
        xcrossy = gy + xdy;
        ycrossx = gx + ydx;
        ycrossy = gy + ydy;
        distX = fabs(xdx) + fabs(xdy);
        distY = fabs(ydx) + fabs(ydy);
      }
      do {
        if (distX < distY) {
          gx = xcrossx;
          gy = xcrossy;
          n = __double2int_rz((double)n - distX); // int -= double, world.cc:931
          // xjumpx = sx*32, xjumpy = sx*32*tana (a power-of-two scaling, exact), world.cc:795-796
          const double xjumpy = ((flags & kXNeg) ? -(double)kRegionWidth : (double)kRegionWidth) * tana;
          xcrossx += (flags & kXNeg) ? -(double)kRegionWidth : (double)kRegionWidth;
          xcrossy += xjumpy;
          distY -= distX;
          distX = (double)kRegionWidth + fabs(xjumpy); // xjumpdist, world.cc:801
        } else {
          gx = ycrossx;
          gy = ycrossy;
          n = __double2int_rz((double)n - distY);
          ycrossx += yjumpx;
          ycrossy += (flags & kYNeg) ? -(double)kRegionWidth : (double)kRegionWidth;
          distX -= distY;
          distY = fabs(yjumpx) + (double)kRegionWidth; // yjumpdist, world.cc:802
        }
        if (n <= 0)
          break;
        if (kSteps)
          region_probes++;
        count = probe_region(g, gx, gy, ix0, iy0, region);
      } while (!count);
      if (n <= 0)
        break;
    }
    {
      // ---- populated region: integer walk over the occupancy masks (world.cc:823-882) ----
      // Both in-region coordinates are counted in the direction of travel: ap along the run axis
      // (bit ap of the direction-normalised mask), cp along the cross axis (mask number cp ^ cflip).
      const int32_t cx0 = ix0 & (kRegionWidth - 1), cy0 = iy0 & (kRegionWidth - 1);
      const int32_t cp0 = (int32_t)((uint32_t)(ymajor ? cx0 : cy0) ^ cflip);
      const int32_t ap0 = (int32_t)((uint32_t)(ymajor ? cy0 : cx0) ^ aflip);
      int32_t cp = cp0, ap = ap0;
      int32_t mdl = -1;
      const bool own = __ldg(&g.head[region].owner) == finder_tag; // (the probe brought its sector in)
      bool walk_it = true;
      if (n >= 2 * kRegionWidth && (flags & kClosedForms)) {
        // The ray does not end in this region (a region holds at most 2 * 32 - 1 of its cells): where and in which state
        // the cell-by-cell walk leaves it, if nothing stops the ray, is a closed form (rt_walk.h).
        int32_t xa, xc, xe;
        walk::region_exit(ap0, cp0, E, b_run, b_cross, xa, xc, xe);
        // Nothing can stop it in a region that holds only the finder's own tree, nor where the occupied cells lie off
        // its path: the tile's summary words say which rows and which columns hold anything at all, and a cell on the
        // path can only be occupied if its row and its column both do.
        bool leave = own;
        if (!own) {
          const uint2 sum = __ldg(reinterpret_cast<const uint2 *>(layer1 ? g.head[region].sum[1] : g.head[region].sum[0]));
          uint32_t rs = ymajor ? sum.x : sum.y, cs = ymajor ? sum.y : sum.x; // along the run axis / the cross axis
          rs ^= (rs ^ __brev(rs)) & revmask;
          cs = cflip ? __brev(cs) : cs;
          rs &= run_mask(ap0, min(xa, kRegionWidth - 1) - ap0 + 1); // the coordinates the walk stands in inside the region
          cs &= run_mask(cp0, min(xc, kRegionWidth - 1) - cp0 + 1);
          leave = !rs || !cs;
          if (!leave) {
            // no cell before the first such row AND the first such column can stop the ray: start the walk where it
            // first stands in both (the later of the two states, each one division away)
            const int32_t a1 = __ffs(rs) - 1, c1 = __ffs(cs) - 1;
            if (c1 > cp0)
              walk::jump_cross(ap0, cp0, E, b_run, b_cross, c1 - cp0, ap, cp, E);
            if (a1 > ap)
              walk::jump_run(ap, cp, E, b_run, b_cross, a1 - ap, ap, cp, E);
            if (kSteps)
              cell_visits += (uint32_t)((ap - ap0) + (cp - cp0));
          }
        }
        if (leave) {
          ap = xa;
          cp = xc;
          E = xe;
          const int32_t steps = (ap - ap0) + (cp - cp0); // one cell visited per step taken
          if (kSteps)
            cell_visits += (uint32_t)steps;
          n -= steps;
          walk_it = false;
        }
      }
      if (walk_it) {
        // the region's row masks (x-major rays) or column masks (y-major); mask number (cp ^ cflip) belongs to cp
        uint32_t tile = occ_layer + region * kTileWords + (ymajor ? kRegionWidth : 0);
        asm volatile("" : "+r"(tile)); // one register to keep, not five instructions to redo in every run
        // no masks are read in a region that holds only the finder's own tree: nothing there can stop this ray
        const int32_t cp_ahead = own ? 0 : kRegionWidth - 1; // masks are requested while cp < this
        uint32_t word = 0;
        if (cp_ahead)
          word = __ldg(g.occ[0] + (tile | ((uint32_t)cp ^ cflip)));
        word ^= (word ^ __brev(word)) & revmask;
        // r = run steps before the next cross step: the smallest r >= 0 with E + r*b_run >= 0. E >= -b_cross
        // always, so r <= r_max. Entering a region E is anywhere in that range (one estimate + fix-up);
        // right after a cross step r is r_max or r_max - 1 (one compare, below).
        int32_t r = 0;
        if (E < 0) {
          const int32_t t = -E;
          r = r_max;
          if (v1 >= t) { // 1 <= ceil(t / b_run) <= r_max - 1 <= 32 and b_run > 0
            r = __float2int_rz((float)t * __frcp_rn((float)b_run));
            while (r * b_run < t)
              r++;
            while (r > 0 && (r - 1) * b_run >= t)
              r--;
          }
        }
        // (n counts the cells the ray may still visit; a region holds at most 2 * 32 - 1 of them, so a ray with
        // more than that left walks the region without looking at n, which is brought up to date afterwards)
        WalkConst wc;
        wc.b_run = b_run; wc.b_cross = b_cross; wc.rm1 = rm1; wc.v1 = v1; wc.cp_ahead = cp_ahead;
        wc.revmask = revmask; wc.aflip = aflip; wc.cflip = cflip; wc.ymajor = ymajor;
        if (n >= 2 * kRegionWidth) {
          mdl = walk_region<kSteps, false>(g, fan, layer1, region, wc, ap, cp, E, n, r, word, tile, cell_visits);
          if (mdl < 0)
            n -= (ap - ap0) + (cp - cp0); // one cell visited per step taken
        } else {
          mdl = walk_region<kSteps, true>(g, fan, layer1, region, wc, ap, cp, E, n, r, word, tile, cell_visits);
        }
      }
      if (mdl >= 0) {
        const uint32_t areal = (uint32_t)ap ^ aflip, creal = (uint32_t)cp ^ cflip;
        hit_model = mdl;
        hit_x = (ix0 & ~(kRegionWidth - 1)) + (int32_t)(ymajor ? creal : areal);
        hit_y = (iy0 & ~(kRegionWidth - 1)) + (int32_t)(ymajor ? areal : creal);
        hit = true;
      }
      // bring the double position up to date (world.cc:868-877 did it step by step)
      const int32_t ka = ap - ap0, kc = cp - cp0; // run / cross steps taken in this region
      gx = seq_add(gx, ymajor ? kc : ka, xneg);
      gy = seq_add(gy, ymajor ? ka : kc, yneg);
      if (hit)
        break;
    }
  }

  res.range = fan->range;
  if (hit) { // world.cc:856-859
    // ax > ay  <=>  x-major with strictly more x than y steps  <=>  !ymajor && b_cross > b_run
    if (!(flags & kYMajor) && b_cross > b_run)
      res.range = fabs((gx - fan->ox * ppm) / trig_major) / ppm;
    else
      res.range = fabs((gy - fan->oy * ppm) / trig_major) / ppm;
  }
  res.hit_model = hit_model;
  res.hit_x = hit_x;
  res.hit_y = hit_y;
  res.cell_visits = cell_visits;
  res.region_probes = region_probes;
}

// ---- the same ray as a resumable state: set-up, one region step at a time, the hit range -------------------------
// For the march kernel that hands a lane its next beam as soon as its ray has ended (rt_march.cu): what trace_ray keeps
// in registers between two iterations of its outer loop, and that loop's body as a function. Same arithmetic, same
// order of operations; only who executes it when differs.
struct RayState {
  double gx, gy, tana, trig_major, yjumpx;
  int32_t n, E, b_run, b_cross, rm1, v1;
  uint32_t flags;      // kYMajor ... kYNeg, | kLayer1
  uint32_t finder_tag;
  const FanRec *fan;
  int32_t hit_x, hit_y; // global cell of the hit (valid once ray_step reports one)
};
enum : uint32_t { kLayer1 = 32u };

__device__ __forceinline__ void ray_setup(const GridView &g, const FanRec *fan, double cosa, double sina, RayState &s)
{
  const double ppm = g.ppm;
  s.fan = fan;
  s.gx = fan->ox * ppm; // world.cc:765-766
  s.gy = fan->oy * ppm;
  s.tana = sina / cosa;
  const double range = fan->range;
  const double dx = ppm * range * cosa, dy = ppm * range * sina; // world.cc:780-781
  const int32_t ax = __double2int_rz(fabs(dx)), ay = __double2int_rz(fabs(dy));
  s.n = ax + ay; // world.cc:792
  const bool ymajor = ay > ax;
  s.b_run = ymajor ? 2 * ax : 2 * ay;
  s.b_cross = ymajor ? 2 * ay : 2 * ax;
  s.E = ymajor ? -(ay - ax) - 1 : (ay - ax);
  const bool xneg = dx < 0, yneg = dy < 0;
  s.flags = (ymajor ? kYMajor : 0u) | ((ymajor ? yneg : xneg) ? kRunNeg : 0u) | ((ymajor ? xneg : yneg) ? kCrossNeg : 0u) |
            (xneg ? kXNeg : 0u) | (yneg ? kYNeg : 0u) | ((fan->layer & 1) ? kLayer1 : 0u);
  int32_t r_max = 33;
  if (s.b_run > 0 && (long long)s.b_run * 33 > s.b_cross)
    r_max = (s.b_cross + s.b_run - 1) / s.b_run;
  s.rm1 = r_max - 1;
  s.v1 = (r_max - 1) * s.b_run;
  s.trig_major = ax > ay ? cosa : sina; // world.cc:856-859
  s.yjumpx = (yneg ? -(double)kRegionWidth : (double)kRegionWidth) / s.tana; // world.cc:797
  s.finder_tag = fan->pred == 2u ? kOwnerNever : __ldg(g.mdl_root + fan->finder) + 1u;
}

// One iteration of World::Raytrace's outer loop (world.cc:818-957): a stretch of empty regions, if the ray stands in one,
// then the walk of the populated region it reaches. Returns true when the ray has ended: hit_model >= 0, or -1 (out of range).
__device__ __forceinline__ bool ray_step(const GridView &g, RayState &s, int32_t &hit_model)
{
  const uint32_t flags = s.flags;
  const bool ymajor = (flags & kYMajor) != 0, runneg = (flags & kRunNeg) != 0;
  const bool xneg = (flags & kXNeg) != 0, yneg = (flags & kYNeg) != 0, layer1 = (flags & kLayer1) != 0;
  const uint32_t aflip = runneg ? kRegionWidth - 1 : 0u, cflip = (flags & kCrossNeg) ? kRegionWidth - 1 : 0u;
  const uint32_t revmask = runneg ? ~0u : 0u;
  const uint32_t occ_layer = layer1 ? (uint32_t)(g.occ[1] - g.occ[0]) : 0u;
  double gx = s.gx, gy = s.gy;
  int32_t n = s.n, E = s.E;
  hit_model = -1;
  int32_t ix0, iy0;
  uint32_t region;
  uint32_t count = probe_region(g, gx, gy, ix0, iy0, region);
  if (!count) {
    const double tana = s.tana;
    double xcrossx, xcrossy, ycrossx, ycrossy, distX, distY;
    {
      double regionx = (double)(ix0 / kRegionWidth * kRegionWidth);
      double regiony = (double)(iy0 / kRegionWidth * kRegionWidth);
      if ((gx < 0) && (ix0 % kRegionWidth))
        regionx -= kRegionWidth;
      if ((gy < 0) && (iy0 % kRegionWidth))
        regiony -= kRegionWidth;
      const double xdx = xneg ? regionx - gx - 1.0 : regionx + kRegionWidth - gx;
      const double xdy = xdx * tana;
      const double ydy = yneg ? regiony - gy - 1.0 : regiony + kRegionWidth - gy;
      const double ydx = ydy / tana;
      xcrossx = gx + xdx;
      xcrossy = gy + xdy;
      ycrossx = gx + ydx;
      ycrossy = gy + ydy;
      distX = fabs(xdx) + fabs(xdy);
      distY = fabs(ydx) + fabs(ydy);
    }
    do {
      if (distX < distY) {
        gx = xcrossx;
        gy = xcrossy;
        n = __double2int_rz((double)n - distX); // int -= double, world.cc:931
        const double xjumpy = (xneg ? -(double)kRegionWidth : (double)kRegionWidth) * tana;
        xcrossx += xneg ? -(double)kRegionWidth : (double)kRegionWidth;
        xcrossy += xjumpy;
        distY -= distX;
        distX = (double)kRegionWidth + fabs(xjumpy);
      } else {
        gx = ycrossx;
        gy = ycrossy;
        n = __double2int_rz((double)n - distY);
        ycrossx += s.yjumpx;
        ycrossy += yneg ? -(double)kRegionWidth : (double)kRegionWidth;
        distX -= distY;
        distY = fabs(s.yjumpx) + (double)kRegionWidth;
      }
      if (n <= 0)
        break;
      count = probe_region(g, gx, gy, ix0, iy0, region);
    } while (!count);
    if (n <= 0) {
      s.gx = gx;
      s.gy = gy;
      s.n = n;
      return true;
    }
  }
  // populated region (world.cc:823-882)
  const int32_t cx0 = ix0 & (kRegionWidth - 1), cy0 = iy0 & (kRegionWidth - 1);
  const int32_t cp0 = (int32_t)((uint32_t)(ymajor ? cx0 : cy0) ^ cflip);
  const int32_t ap0 = (int32_t)((uint32_t)(ymajor ? cy0 : cx0) ^ aflip);
  int32_t cp = cp0, ap = ap0;
  uint32_t tile = occ_layer + region * kTileWords + (ymajor ? kRegionWidth : 0);
  const int32_t cp_ahead = __ldg(&g.head[region].owner) != s.finder_tag ? kRegionWidth - 1 : 0;
  uint32_t word = 0;
  if (cp_ahead)
    word = __ldg(g.occ[0] + (tile | ((uint32_t)cp ^ cflip)));
  word ^= (word ^ __brev(word)) & revmask;
  int32_t r = 0;
  if (E < 0) {
    const int32_t t = -E;
    r = s.rm1 + 1;
    if (s.v1 >= t) {
      r = __float2int_rz((float)t * __frcp_rn((float)s.b_run));
      while (r * s.b_run < t)
        r++;
      while (r > 0 && (r - 1) * s.b_run >= t)
        r--;
    }
  }
  WalkConst wc;
  wc.b_run = s.b_run; wc.b_cross = s.b_cross; wc.rm1 = s.rm1; wc.v1 = s.v1; wc.cp_ahead = cp_ahead;
  wc.revmask = revmask; wc.aflip = aflip; wc.cflip = cflip; wc.ymajor = ymajor;
  uint32_t unused = 0;
  int32_t mdl;
  if (n >= 2 * kRegionWidth) {
    mdl = walk_region<false, false>(g, s.fan, layer1, region, wc, ap, cp, E, n, r, word, tile, unused);
    if (mdl < 0)
      n -= (ap - ap0) + (cp - cp0);
  } else {
    mdl = walk_region<false, true>(g, s.fan, layer1, region, wc, ap, cp, E, n, r, word, tile, unused);
  }
  if (mdl >= 0) {
    const uint32_t areal = (uint32_t)ap ^ aflip, creal = (uint32_t)cp ^ cflip;
    s.hit_x = (ix0 & ~(kRegionWidth - 1)) + (int32_t)(ymajor ? creal : areal);
    s.hit_y = (iy0 & ~(kRegionWidth - 1)) + (int32_t)(ymajor ? areal : creal);
  }
  const int32_t ka = ap - ap0, kc = cp - cp0;
  s.gx = seq_add(gx, ymajor ? kc : ka, xneg);
  s.gy = seq_add(gy, ymajor ? ka : kc, yneg);
  s.n = n;
  s.E = E;
  hit_model = mdl;
  return mdl >= 0 || n <= 0;
}

// RaytraceResult::range (world.cc:856-859; the ray's own range when nothing was hit)
__device__ __forceinline__ double ray_range(const GridView &g, const RayState &s, int32_t hit_model)
{
  if (hit_model < 0)
    return s.fan->range;
  if (!(s.flags & kYMajor) && s.b_cross > s.b_run)
    return fabs((s.gx - s.fan->ox * g.ppm) / s.trig_major) / g.ppm;
  return fabs((s.gy - s.fan->oy * g.ppm) / s.trig_major) / g.ppm;
}

// cos / sin of a heading as World::Raytrace takes them (world.cc:773-775): the host C library's
// values, bit for bit (rt_libm.cuh), because the traversal is chaotic in their last bit.
__device__ __forceinline__ void heading_trig(double heading, double &cosa, double &sina)
{
  const double ang = heading == 0.0 ? 1e-12 : heading;
  if (libm::covered(ang)) {
    sina = libm::sin_host_exact(ang);
    cosa = libm::cos_host_exact(ang);
  } else { // |heading| > 1e8 or not finite: no sensor produces this
    sina = sin(ang);
    cosa = cos(ang);
  }
}

} // namespace trace
} // namespace stg
