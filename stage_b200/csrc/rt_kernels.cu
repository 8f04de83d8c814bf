// sm_100a kernels of the Stage sensor-raycast path.
//   K1  apply_deltas_*   block rasterisation: per-tick moved-block deltas -> occupancy tiles,
//                        region counts and the cell-content table
//                        (World::MapPoly world.cc:990-1056, Cell::AddBlock/RemoveBlock
//                        region.cc:283-299, Block::UnMap block.cc:183-189)
//   K2a fan_headings     per-beam headings (ModelRanger::Sensor::Update model_ranger.cc:232-255,
//                        World::Raytrace(fan) world.cc:731-743)
//   K2  ray_march        batched DDA march, one lane per beam, predicates in-kernel
//                        (World::Raytrace world.cc:756-963) with the K3 ranger epilogue fused
//                        (intensity / bearing, model_ranger.cc:243-251)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (the traversal's FP64 address
// math must round exactly like the reference's x86-64 -O2 build, which never contracts a*b+c).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_device.h"
#include "rt_libm.cuh"

namespace stg {

// ------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned long long ld_slot(const unsigned long long *p)
{
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Cell visit k (0-based) of the integer line (x0,y0)->(x1,y1) walked as World::MapPoly does
// (world.cc:1000-1052): error term exy starts at ay-ax; exy<0 -> x step, exy+=2ay; else y step,
// exy-=2ax. exy stays in [-2ax, 2ay), so the number of x steps among the first k steps is the
// unique i with -2ax <= (ay-ax) - 2ax*k + i*(2ax+2ay) < 2ay, i.e. floor((2ax*k+ax+ay-1)/(2ax+2ay)).
__device__ __forceinline__ void edge_cell(const EdgeRec &e, uint32_t k, int32_t &x, int32_t &y)
{
  const int32_t dx = e.x1 - e.x0, dy = e.y1 - e.y0;
  const int32_t sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
  const long long ax = dx < 0 ? -(long long)dx : dx, ay = dy < 0 ? -(long long)dy : dy;
  const long long i = (2 * ax * (long long)k + ax + ay - 1) / (2 * (ax + ay));
  x = e.x0 + sx * (int32_t)i;
  y = e.y0 + sy * (int32_t)((long long)k - i);
}

// Find the edge whose step range contains global step t (first_step is a running sum).
__device__ __forceinline__ uint32_t find_edge(const EdgeRec *edges, uint32_t n, uint32_t t)
{
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (edges[mid].first_step <= t)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

struct CellAddr {
  uint32_t region; // row-major region index inside the extent
  uint32_t key;    // region*1024 + cy*32 + cx
  uint32_t cx, cy;
  bool inside;     // the cell lies in the device extent (the host checks every outline; a record that fails here is corrupt)
};

__device__ __forceinline__ CellAddr cell_addr(const GridView &g, int32_t x, int32_t y)
{
  CellAddr c;
  const int32_t rix = (x >> kRBits) - g.rx0, riy = (y >> kRBits) - g.ry0;
  c.inside = (uint32_t)rix < (uint32_t)g.nrx && (uint32_t)riy < (uint32_t)g.nry;
  c.region = (uint32_t)(riy * g.nrx + rix);
  c.cx = (uint32_t)x & (kRegionWidth - 1);
  c.cy = (uint32_t)y & (kRegionWidth - 1);
  c.key = (c.region << (2 * kRBits)) | (c.cy << kRBits) | c.cx;
  return c;
}

// ------------------------------------------------------------------------------------------
// K1: delta application. One thread per cell visit of every edge of every rank's blob.
// Three phases with a grid-wide order between them (separate launches on one stream):
//   remove : drop (cell, block) entries, decrement Region::count, clear the occupancy bit
//   fixup  : re-set the bit of every cell touched by a removal that still holds another block
//   insert : add entries, increment Region::count, set the bit
// ------------------------------------------------------------------------------------------

struct BlobView {
  const DeltaHeader *hdr;
  const EdgeRec *remove_edges, *insert_edges;
  const BlockUpd *block_upd;
  const ModelUpd *model_upd;
  const PoseUpd *pose_upd;
  uint32_t n_remove, n_insert, n_block_upd, n_model_upd, n_pose_upd; // 0 for a slot that does not hold a well-formed blob
};

// A slot holds a blob only if it says so itself: the magic word, and record counts that add up to its own size and
// fit the slot. Anything else (a gather that ran before the export had landed, a stale or foreign buffer) is treated as
// an empty blob and reported through GridView::err_host, so a torn header can never index past its slot.
__device__ __forceinline__ bool blob_ok(const DeltaHeader *h, size_t slot_bytes)
{
  if (h->magic != kDeltaMagic)
    return false;
  const unsigned long long need = sizeof(DeltaHeader) + ((unsigned long long)h->n_remove + h->n_insert) * sizeof(EdgeRec) +
                                  (unsigned long long)h->n_block_upd * sizeof(BlockUpd) + (unsigned long long)h->n_model_upd * sizeof(ModelUpd) +
                                  (unsigned long long)h->n_pose_upd * sizeof(PoseUpd);
  return need <= slot_bytes && need == h->n_bytes && h->steps_remove == h->steps_remove_l[0] + h->steps_remove_l[1] &&
         h->steps_insert == h->steps_insert_l[0] + h->steps_insert_l[1];
}

__device__ __forceinline__ void raise_error(const GridView &g, uint32_t bit)
{
  if (g.err_host)
    *reinterpret_cast<volatile uint32_t *>(g.err_host) = bit; // a plain store: any non-zero value fails the next stg_rt_sync
}

__device__ __forceinline__ BlobView blob_view(const unsigned char *blobs, int rank, size_t slot_bytes)
{
  BlobView v;
  const unsigned char *p = blobs + (size_t)rank * slot_bytes;
  v.hdr = reinterpret_cast<const DeltaHeader *>(p);
  const bool ok = blob_ok(v.hdr, slot_bytes);
  v.n_remove = ok ? v.hdr->n_remove : 0u;
  v.n_insert = ok ? v.hdr->n_insert : 0u;
  v.n_block_upd = ok ? v.hdr->n_block_upd : 0u;
  v.n_model_upd = ok ? v.hdr->n_model_upd : 0u;
  v.n_pose_upd = ok ? v.hdr->n_pose_upd : 0u;
  v.remove_edges = reinterpret_cast<const EdgeRec *>(p + sizeof(DeltaHeader));
  v.insert_edges = v.remove_edges + v.n_remove;
  v.block_upd = reinterpret_cast<const BlockUpd *>(v.insert_edges + v.n_insert);
  v.model_upd = reinterpret_cast<const ModelUpd *>(v.block_upd + v.n_block_upd);
  v.pose_upd = reinterpret_cast<const PoseUpd *>(v.model_upd + v.n_model_upd);
  return v;
}

// ---- what one cell visit does to the grid (shared by the edge-record path and the pose path) ------------------
// The occupancy bits of a cell in its region's tile, and the tile's summary words (GridView::occ_sum: which row masks
// and which column masks are non-zero). The summaries only ever have to be a SUPERSET of the truth for the march to be
// right (a bit too many costs it a walk it could have skipped), so K1 keeps them that way at no cost to its dependent
// chains: setting a tile bit sets the two summary bits with fire-and-forget atomics, clearing one leaves them standing
// and marks the region in GridView::occ_dirty; k1_summaries recomputes the marked regions' words from their tiles after
// the pass, on a stream of its own, while the march is already running (it may read either version of a word).
// (Tried first: whoever clears the last bit of a mask clears its summary bit, whoever sets the first sets it - those
// atomics have to return their old values, and the round trips made a 1000-robot pass 0.024 ms longer; then the
// recomputing kernel in line behind the pass: 0.013 ms.)
__device__ __forceinline__ void occ_clear(const GridView &g, const CellAddr &c, uint32_t L)
{
  uint32_t *tile = g.occ[L] + (size_t)c.region * kTileWords;
  atomicAnd(tile + c.cy, ~(1u << c.cx));
  atomicAnd(tile + kRegionWidth + c.cx, ~(1u << c.cy));
  atomicOr(g.occ_dirty + (c.region >> 5), 1u << (c.region & 31u));
}
__device__ __forceinline__ void occ_set(const GridView &g, const CellAddr &c, uint32_t L)
{
  uint32_t *tile = g.occ[L] + (size_t)c.region * kTileWords;
  uint32_t *sum = g.head[c.region].sum[L];
  atomicOr(tile + c.cy, 1u << c.cx);
  atomicOr(tile + kRegionWidth + c.cx, 1u << c.cy);
  atomicOr(sum, 1u << c.cy);
  atomicOr(sum + 1, 1u << c.cx);
}

// Cell::RemoveBlock (region.cc:292-299, EraseAll: every copy goes) + Region::RemoveBlock's count
__device__ __forceinline__ void cell_remove(const GridView &g, const CellAddr &c, uint32_t block, uint32_t L)
{
  const unsigned long long entry = ((unsigned long long)c.key << 32) | (unsigned long long)(block + 1u);
  unsigned long long *slots = g.slots[L];
  uint32_t h = slot_home(c.key, g.slot_mask);
  for (uint32_t probes = 0; probes <= g.slot_mask; probes++) {
    const unsigned long long cur = ld_slot(slots + h);
    if (cur == kSlotEmpty)
      break;
    if (cur == entry)
      atomicCAS(slots + h, entry, kSlotTomb);
    h = (h + 1) & g.slot_mask;
  }
  if (atomicSub(&g.head[c.region].count, 1u) == 1u) // the region's last entry: it belongs to nobody again
    g.head[c.region].owner = kOwnerNone;
  occ_clear(g, c, L);
}

// after every removal has landed: a cell that lost an entry but still holds another block gets its mask bits back
__device__ __forceinline__ void cell_fixup(const GridView &g, const CellAddr &c, uint32_t L)
{
  const unsigned long long *slots = g.slots[L];
  uint32_t h = slot_home(c.key, g.slot_mask);
  for (uint32_t probes = 0; probes <= g.slot_mask; probes++) {
    const unsigned long long cur = ld_slot(slots + h);
    if (cur == kSlotEmpty)
      break;
    if ((uint32_t)(cur >> 32) == c.key) { // tombstones carry key 0xffffffff, never a cell key
      occ_set(g, c, L);
      break;
    }
    h = (h + 1) & g.slot_mask;
  }
}

// Cell::AddBlock (region.cc:283-290) + Region::AddBlock's count, the occupancy bits and the region's owner tag
__device__ __forceinline__ void cell_insert(const GridView &g, const CellAddr &c, uint32_t block, uint32_t L)
{
  const unsigned long long entry = ((unsigned long long)c.key << 32) | (unsigned long long)(block + 1u);
  unsigned long long *slots = g.slots[L];
  uint32_t h = slot_home(c.key, g.slot_mask);
  bool placed = false;
  for (uint32_t probes = 0; probes <= g.slot_mask;) {
    const unsigned long long cur = ld_slot(slots + h);
    if (cur == kSlotEmpty || cur == kSlotTomb) {
      if (atomicCAS(slots + h, cur, entry) == cur) {
        placed = true;
        break;
      }
      continue; // lost the race for this slot: look at it again
    }
    h = (h + 1) & g.slot_mask;
    probes++;
  }
  if (!placed) { // every slot taken: no entry, so no count and no mask bit either (a bit nothing resolves would make
    raise_error(g, kErrTableFull); // the cell transparent and the region count could never return to zero)
    return;
  }
  atomicAdd(&g.head[c.region].count, 1u);
  { // region owner: stays a tree's tag only while every entry added belongs to that tree
    const uint32_t tag = (g.blk_rec[0][block].root_flags & kRecRootMask) + 1u;
    uint32_t *owner = &g.head[c.region].owner;
    uint32_t cur = *reinterpret_cast<volatile uint32_t *>(owner);
    if (cur == kOwnerNone)
      cur = atomicCAS(owner, kOwnerNone, tag);
    if (cur != kOwnerNone && cur != tag && cur != kOwnerMixed)
      atomicExch(owner, kOwnerMixed);
  }
  occ_set(g, c, L);
}

// The cell visits of all ranks' blobs as ONE index space, so that a gathered delta (N ranks) is a
// single grid-stride pass instead of N latency-bound passes. prefix[] (shared, n_ranks + 1 entries)
// holds the running sums of the ranks' step counts; ranks beyond kMaxRanksFlat fall back to rank-by-
// rank passes. The order in which visits of different ranks land does not matter: removals are
// independent, insertions claim slots by CAS, and the order of a cell's list is carried by seq.
constexpr int kMaxRanksFlat = 64;

template <bool kInsert>
__device__ __forceinline__ void k1_step_prefix(const GridView &g, const unsigned char *blobs, int n_ranks, size_t slot_bytes, uint32_t *prefix)
{
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int r = 0; r < n_ranks; r++) {
      prefix[r] = run;
      const DeltaHeader *h = reinterpret_cast<const DeltaHeader *>(blobs + (size_t)r * slot_bytes);
      if (blob_ok(h, slot_bytes))
        run += kInsert ? h->steps_insert : h->steps_remove;
      else if (blockIdx.x == 0)
        raise_error(g, kErrBadBlob);
    }
    prefix[n_ranks] = run;
  }
  __syncthreads();
}

__device__ __forceinline__ int k1_rank_of(const uint32_t *prefix, int n_ranks, uint32_t t)
{
  int r = 0;
  while (r + 1 < n_ranks && prefix[r + 1] <= t)
    r++;
  return r;
}

// The epoch of a pass: the largest one any of its blobs proposes (every sender proposes its own next epoch; replicas
// that exchanged the same blobs therefore stamp the same sequence numbers, whatever each did locally before).
__device__ __forceinline__ unsigned long long k1_pass_epoch(const unsigned char *blobs, int n_ranks, size_t slot_bytes, unsigned long long epoch)
{
  for (int r = 0; r < n_ranks; r++) {
    const DeltaHeader *h = reinterpret_cast<const DeltaHeader *>(blobs + (size_t)r * slot_bytes);
    if (blob_ok(h, slot_bytes) && h->epoch > epoch)
      epoch = h->epoch;
  }
  return epoch;
}

__device__ __forceinline__ void k1_table_updates(const GridView &g, const unsigned char *blobs, int n_ranks,
                                                 size_t slot_bytes, unsigned long long epoch)
{
  epoch = k1_pass_epoch(blobs, n_ranks, slot_bytes, epoch);
  for (int rank = 0; rank < n_ranks; rank++) {
    const BlobView v = blob_view(blobs, rank, slot_bytes);
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = t0; i < v.n_model_upd; i += stride) {
      const ModelUpd u = v.model_upd[i];
      if (u.id < g.n_models) {
        g.mdl_root[u.id] = u.root;
        g.mdl_ranger_return[u.id] = u.ranger_return;
        g.mdl_flags[u.id] = u.flags;
        g.mdl_pal[u.id] = (uint8_t)u.pal;
      }
    }
    for (uint32_t i = t0; i < v.n_block_upd; i += stride) {
      const BlockUpd u = v.block_upd[i];
      if (u.block < g.n_blocks) {
        // A block rendered before this pass changes trees (its model was re-parented, or the id now names a
        // block of another model): what the region owners say about its cells no longer holds. "Before this
        // pass" = it carries a sequence number of an earlier epoch; one written by this very launch (the
        // block's own Map record may be applied by another thread right now) does not count.
        if ((g.blk_rec[0][u.block].root_flags ^ u.root_flags) & kRecRootMask) {
          const unsigned long long s0 = g.blk_rec[0][u.block].seq, s1 = g.blk_rec[1][u.block].seq;
          if ((s0 && (s0 >> kSeqEpochShift) != epoch) || (s1 && (s1 >> kSeqEpochShift) != epoch))
            g.head[g.n_regions].owner = 1u;
        }
        for (int L = 0; L < 2; L++) { // z bounds and ownership are per block: both layers see them
          BlockRec *rec = g.blk_rec[L] + u.block;
          rec->zmin = u.zmin;
          rec->zmax = u.zmax;
          rec->model = u.model;
          rec->root_flags = u.root_flags;
        }
        // insertion order across the whole job: epoch, then rank, then call order within the rank
        if (!(u.layer & kUpdKeepSeq))
          g.blk_rec[u.layer & 1][u.block].seq = (epoch << kSeqEpochShift) |
                                                ((unsigned long long)(v.hdr->rank & 0xff) << kSeqRankShift) | (u.seq_local & kSeqLocalMax);
      }
    }
  }
}

__device__ __forceinline__ void k1_remove(const GridView &g, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                                          const uint32_t *prefix)
{
  {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t tt = blockIdx.x * blockDim.x + threadIdx.x; tt < prefix[n_ranks]; tt += stride) {
      const int rank = k1_rank_of(prefix, n_ranks, tt);
      const BlobView v = blob_view(blobs, rank, slot_bytes);
      const uint32_t t = tt - prefix[rank];
      const EdgeRec e = v.remove_edges[find_edge(v.remove_edges, v.n_remove, t)];
      int32_t x, y;
      edge_cell(e, t - e.first_step, x, y);
      const CellAddr c = cell_addr(g, x, y);
      if (!c.inside || e.block >= g.n_blocks) {
        raise_error(g, kErrOutOfExtent);
        continue;
      }
      cell_remove(g, c, e.block, e.layer & 1);
    }
  }
}

__device__ __forceinline__ void k1_fixup(const GridView &g, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                                         const uint32_t *prefix)
{
  {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t tt = blockIdx.x * blockDim.x + threadIdx.x; tt < prefix[n_ranks]; tt += stride) {
      const int rank = k1_rank_of(prefix, n_ranks, tt);
      const BlobView v = blob_view(blobs, rank, slot_bytes);
      const uint32_t t = tt - prefix[rank];
      const EdgeRec e = v.remove_edges[find_edge(v.remove_edges, v.n_remove, t)];
      int32_t x, y;
      edge_cell(e, t - e.first_step, x, y);
      const CellAddr c = cell_addr(g, x, y);
      if (!c.inside || e.block >= g.n_blocks) {
        raise_error(g, kErrOutOfExtent);
        continue;
      }
      cell_fixup(g, c, e.layer & 1);
    }
  }
}

__device__ __forceinline__ void k1_insert(const GridView &g, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                                          const uint32_t *prefix)
{
  {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t tt = blockIdx.x * blockDim.x + threadIdx.x; tt < prefix[n_ranks]; tt += stride) {
      const int rank = k1_rank_of(prefix, n_ranks, tt);
      const BlobView v = blob_view(blobs, rank, slot_bytes);
      const uint32_t t = tt - prefix[rank];
      const EdgeRec e = v.insert_edges[find_edge(v.insert_edges, v.n_insert, t)];
      int32_t x, y;
      edge_cell(e, t - e.first_step, x, y);
      const CellAddr c = cell_addr(g, x, y);
      if (!c.inside || e.block >= g.n_blocks) {
        raise_error(g, kErrOutOfExtent);
        continue;
      }
      cell_insert(g, c, e.block, e.layer & 1);
    }
  }
}

// phase 0: table updates and removals (independent of each other).
__global__ void __launch_bounds__(256) k1_phase0(GridView g, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                                                 unsigned long long epoch)
{
  __shared__ uint32_t prefix[kMaxRanksFlat + 1];
  k1_table_updates(g, blobs, n_ranks, slot_bytes, epoch);
  for (int r0 = 0; r0 < n_ranks; r0 += kMaxRanksFlat) { // one pass unless there are more ranks than that
    const int nr = min(n_ranks - r0, kMaxRanksFlat);
    const unsigned char *part = blobs + (size_t)r0 * slot_bytes;
    k1_step_prefix<false>(g, part, nr, slot_bytes, prefix);
    k1_remove(g, part, nr, slot_bytes, prefix);
    __syncthreads();
  }
}

// phase 1, after every removal has landed: mask fix-up of the removed cells and the insertions. The
// two only ever SET mask bits, and a fix-up that already sees a freshly inserted entry sets a bit the
// insertion sets anyway, so they can share a launch.
__global__ void __launch_bounds__(256) k1_phase1(GridView g, const unsigned char *blobs, int n_ranks, size_t slot_bytes)
{
  __shared__ uint32_t prefix[kMaxRanksFlat + 1];
  for (int r0 = 0; r0 < n_ranks; r0 += kMaxRanksFlat) {
    const int nr = min(n_ranks - r0, kMaxRanksFlat);
    const unsigned char *part = blobs + (size_t)r0 * slot_bytes;
    k1_step_prefix<false>(g, part, nr, slot_bytes, prefix);
    k1_fixup(g, part, nr, slot_bytes, prefix);
    __syncthreads();
    k1_step_prefix<true>(g, part, nr, slot_bytes, prefix);
    k1_insert(g, part, nr, slot_bytes, prefix);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// K1, pose path: Block::UnMap + Block::Map of blocks whose outline the device derives itself from the model's pose
// (Model::LocalToPixels, model.cc:474-491) and the block's registered local polygon. One warp per PoseUpd record, all
// ranks' records as one index space; the lanes share the cell visits of each edge of the outline. Same two phases as
// the edge-record path, launched right behind them:
//   phase 0: the new pixel outline -> GeomView::pix_new; the block's BlockRec (global_z = local_z + pose z,
//            block.cc:177-180; owner; sequence number); removal of the outline rendered before (GeomView::pix)
//   phase 1: mask fix-up over the old outline, insertion of the new one, pix <- pix_new
// ------------------------------------------------------------------------------------------

// prefix[r] = pose records of the ranks before r; *est_steps = the senders' estimate of the cell visits of the outlines those
// records take out or put in, whichever is more (DeltaHeader::est_remove_l / est_insert_l)
__device__ __forceinline__ void k1_pose_prefix(const GridView &g, const unsigned char *blobs, int n_ranks, size_t slot_bytes, uint32_t *prefix,
                                               uint32_t *est_steps)
{
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    unsigned long long out = 0, in = 0;
    for (int r = 0; r < n_ranks; r++) {
      prefix[r] = run;
      const DeltaHeader *h = reinterpret_cast<const DeltaHeader *>(blobs + (size_t)r * slot_bytes);
      if (blob_ok(h, slot_bytes)) {
        run += h->n_pose_upd;
        out += (unsigned long long)h->est_remove_l[0] + h->est_remove_l[1];
        in += (unsigned long long)h->est_insert_l[0] + h->est_insert_l[1];
      }
    }
    prefix[n_ranks] = run;
    *est_steps = (uint32_t)min(max(out, in), 0xffffffffull);
  }
  __syncthreads();
}

// the cell visits of a closed pixel outline, edge after edge as World::MapPoly walks them (world.cc:995-1052), shared
// out over the lanes of the warp. Outlines of up to 32 vertices (every robot body) are one index space: lane i holds
// edge i's step count, a warp scan turns them into offsets, and visit s of the outline goes to lane s % 32 - a 0.1 m
// puck (20 visits) or a 0.44 x 0.38 m robot (82) is done in one to three rounds instead of one round per edge.
// part / n_parts: the rounds of 32 visits are dealt out to n_parts warps that share the outline (round r goes to warp
// r % n_parts), so that a robot's three rounds are three dependent chains side by side instead of one after the other.
template <class Op>
__device__ __forceinline__ void outline_cells(const GridView &g, const int32_t *pix, uint32_t n_pts, uint32_t lane, Op op, uint32_t part = 0,
                                              uint32_t n_parts = 1)
{
  if (n_pts <= 32) {
    int32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    uint32_t steps = 0;
    if (lane < n_pts) {
      const uint32_t j = lane + 1 == n_pts ? 0 : lane + 1;
      x0 = pix[2 * lane];
      y0 = pix[2 * lane + 1];
      x1 = pix[2 * j];
      y1 = pix[2 * j + 1];
      const long long adx = x1 > x0 ? (long long)x1 - x0 : (long long)x0 - x1;
      const long long ady = y1 > y0 ? (long long)y1 - y0 : (long long)y0 - y1;
      steps = (uint32_t)(adx + ady);
    }
    uint32_t incl = steps; // inclusive scan of the step counts
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d)
        incl += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    for (uint32_t s0 = 32 * part; s0 < total; s0 += 32 * n_parts) {
      const uint32_t s = s0 + lane;
      // which edge holds visit s: the first lane whose inclusive count exceeds it
      uint32_t e = 0;
      for (uint32_t i = 0; i < n_pts; i++) { // n_pts is warp-uniform: every lane takes part in every shuffle
        const uint32_t inc_i = __shfl_sync(0xffffffffu, incl, i);
        e += (s < total && inc_i <= s) ? 1u : 0u;
      }
      EdgeRec ed;
      ed.x0 = __shfl_sync(0xffffffffu, x0, e & 31u);
      ed.y0 = __shfl_sync(0xffffffffu, y0, e & 31u);
      ed.x1 = __shfl_sync(0xffffffffu, x1, e & 31u);
      ed.y1 = __shfl_sync(0xffffffffu, y1, e & 31u);
      const uint32_t inc_e = __shfl_sync(0xffffffffu, incl, e & 31u), st_e = __shfl_sync(0xffffffffu, steps, e & 31u);
      if (s < total) {
        int32_t x, y;
        edge_cell(ed, s - (inc_e - st_e), x, y);
        const CellAddr c = cell_addr(g, x, y);
        if (!c.inside)
          raise_error(g, kErrOutOfExtent);
        else
          op(c);
      }
    }
    return;
  }
  for (uint32_t i = part; i < n_pts; i += n_parts) { // long outlines (a traced floorplan): edge by edge
    const uint32_t j = i + 1 == n_pts ? 0 : i + 1;
    EdgeRec e;
    e.x0 = pix[2 * i];
    e.y0 = pix[2 * i + 1];
    e.x1 = pix[2 * j];
    e.y1 = pix[2 * j + 1];
    const long long adx = e.x1 > e.x0 ? (long long)e.x1 - e.x0 : (long long)e.x0 - e.x1;
    const long long ady = e.y1 > e.y0 ? (long long)e.y1 - e.y0 : (long long)e.y0 - e.y1;
    const uint32_t n_steps = (uint32_t)(adx + ady);
    for (uint32_t k = lane; k < n_steps; k += 32) {
      int32_t x, y;
      edge_cell(e, k, x, y);
      const CellAddr c = cell_addr(g, x, y);
      if (!c.inside) {
        raise_error(g, kErrOutOfExtent);
        continue;
      }
      op(c);
    }
  }
}

// Warps that share the cell visits of one pose record's outline (82 visits for a 0.44 x 0.38 m robot: three rounds of 32). A
// pass over a few thousand records is all latency and leaves most of the device idle: three chains side by side shorten it
// (1000 robots: 0.077 -> 0.054 ms). Bodies of one round gain nothing and the extra warps get in the way (100,000 pucks:
// 0.87 -> 1.42 ms with three), so they keep one.
// Decided by what the records hold (the senders' estimate of their outlines' cell visits, DeltaHeader) rather than by how
// many there are: three warps where an outline is two and more rounds on average (8000 robots: 0.267 -> 0.249 ms).
__device__ __forceinline__ uint32_t pose_parts(uint32_t n_records, uint32_t est_steps)
{
  return n_records && est_steps / n_records >= 64u ? 3u : 1u;
}
struct OpRemove {
  const GridView &g;
  uint32_t block, L;
  __device__ void operator()(const CellAddr &c) const { cell_remove(g, c, block, L); }
};
struct OpFixup {
  const GridView &g;
  uint32_t L;
  __device__ void operator()(const CellAddr &c) const { cell_fixup(g, c, L); }
};
struct OpInsert {
  const GridView &g;
  uint32_t block, L;
  __device__ void operator()(const CellAddr &c) const { cell_insert(g, c, block, L); }
};

__global__ void __launch_bounds__(256) k1_pose_phase0(GridView g, GeomView gv, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                                                      unsigned long long epoch)
{
  __shared__ uint32_t prefix[kMaxRanksFlat + 1];
  __shared__ uint32_t est_steps;
  const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  epoch = k1_pass_epoch(blobs, n_ranks, slot_bytes, epoch);
  for (int r0 = 0; r0 < n_ranks; r0 += kMaxRanksFlat) {
    const int nr = min(n_ranks - r0, kMaxRanksFlat);
    const unsigned char *part = blobs + (size_t)r0 * slot_bytes;
    k1_pose_prefix(g, part, nr, slot_bytes, prefix, &est_steps);
    // 1 + kPoseParts warps share a record: one derives the new outline and the block's record, the others take the old
    // outline out, a round of 32 cell visits each - dependent chains side by side instead of one after the other (a pass
    // over 1000 robots is all latency)
    const uint32_t kPoseParts = pose_parts(prefix[nr], est_steps);
    for (uint32_t ww = warp; ww < (1 + kPoseParts) * prefix[nr]; ww += n_warps) {
      const uint32_t w = ww / (1 + kPoseParts), task = ww % (1 + kPoseParts);
      const bool second = task != 0;
      const int rank = k1_rank_of(prefix, nr, w);
      const BlobView v = blob_view(part, rank, slot_bytes);
      const PoseUpd u = v.pose_upd[w - prefix[rank]];
      if (u.block >= gv.n_blocks || u.block >= g.n_blocks) {
        if (lane == 0 && !second)
          raise_error(g, kErrOutOfExtent);
        continue;
      }
      const BlockGeom bg = gv.geom[u.block];
      const uint32_t L = u.layer_flags & 1u;
      if (!second && !(u.layer_flags & kPoseNoInsert)) {
        // Model::LocalToPixels: ptpose = gpose + Pose(pt) (Pose::operator+, stage.hh:297-304: cos / sin of gpose.a from the
        // host C library's algorithm, rt_libm.cuh; no contraction), pixel = (int32)floor(coordinate * ppm)
        double cosa, sina;
        if (libm::covered(u.a)) {
          sina = libm::sin_host_exact(u.a);
          cosa = libm::cos_host_exact(u.a);
        } else {
          sina = sin(u.a);
          cosa = cos(u.a);
        }
        for (uint32_t i = lane; i < bg.n_pts; i += 32) {
          const double px = gv.pts[2 * (size_t)(bg.first_pt + i)], py = gv.pts[2 * (size_t)(bg.first_pt + i) + 1];
          const double gx = u.x + px * cosa - py * sina, gy = u.y + px * sina + py * cosa;
          gv.pix_new[L][2 * (size_t)(bg.first_pt + i)] = (int32_t)floor(gx * g.ppm);
          gv.pix_new[L][2 * (size_t)(bg.first_pt + i) + 1] = (int32_t)floor(gy * g.ppm);
        }
        if (lane == 0) { // Block::Map's second half (block.cc:176-180) and the block's place in the cell lists
          if ((g.blk_rec[0][u.block].root_flags ^ u.root_flags) & kRecRootMask) {
            const unsigned long long s0 = g.blk_rec[0][u.block].seq, s1 = g.blk_rec[1][u.block].seq;
            if ((s0 && (s0 >> kSeqEpochShift) != epoch) || (s1 && (s1 >> kSeqEpochShift) != epoch))
              g.head[g.n_regions].owner = 1u;
          }
          for (int l = 0; l < 2; l++) {
            BlockRec *rec = g.blk_rec[l] + u.block;
            rec->zmin = bg.zmin + u.z;
            rec->zmax = bg.zmax + u.z;
            rec->model = bg.model;
            rec->root_flags = u.root_flags;
          }
          g.blk_rec[L][u.block].seq = (epoch << kSeqEpochShift) | ((unsigned long long)(v.hdr->rank & 0xff) << kSeqRankShift) | (u.seq_local & kSeqLocalMax);
        }
      }
      if (second && (u.layer_flags & kPoseRemoveOld) && 32u * (task - 1) < bg.est_steps) { // (a puck's 20 visits are one round)
        const OpRemove op = { g, u.block, L };
        outline_cells(g, gv.pix[L] + 2 * (size_t)bg.first_pt, bg.n_pts, lane, op, task - 1, kPoseParts);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k1_pose_phase1(GridView g, GeomView gv, const unsigned char *blobs, int n_ranks, size_t slot_bytes)
{
  __shared__ uint32_t prefix[kMaxRanksFlat + 1];
  __shared__ uint32_t est_steps;
  const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = 0; r0 < n_ranks; r0 += kMaxRanksFlat) {
    const int nr = min(n_ranks - r0, kMaxRanksFlat);
    const unsigned char *part = blobs + (size_t)r0 * slot_bytes;
    k1_pose_prefix(g, part, nr, slot_bytes, prefix, &est_steps);
    // again 1 + kPoseParts warps per record: kPoseParts insert the new outline, a round of 32 cell visits each, the last one
    // fixes the masks up over the old one and then - it is the only reader of the old outline - puts the new one in its place
    const uint32_t kPoseParts = pose_parts(prefix[nr], est_steps);
    for (uint32_t ww = warp; ww < (1 + kPoseParts) * prefix[nr]; ww += n_warps) {
      const uint32_t w = ww / (1 + kPoseParts), task = ww % (1 + kPoseParts);
      const bool second = task == kPoseParts;
      const int rank = k1_rank_of(prefix, nr, w);
      const BlobView v = blob_view(part, rank, slot_bytes);
      const PoseUpd u = v.pose_upd[w - prefix[rank]];
      if (u.block >= gv.n_blocks || u.block >= g.n_blocks)
        continue;
      const BlockGeom bg = gv.geom[u.block];
      const uint32_t L = u.layer_flags & 1u;
      int32_t *pix = gv.pix[L] + 2 * (size_t)bg.first_pt;
      const int32_t *fresh = gv.pix_new[L] + 2 * (size_t)bg.first_pt;
      if (!second) {
        if (!(u.layer_flags & kPoseNoInsert) && 32u * task < bg.est_steps) {
          const OpInsert op = { g, u.block, L };
          outline_cells(g, fresh, bg.n_pts, lane, op, task, kPoseParts);
        }
        continue;
      }
      if (u.layer_flags & kPoseRemoveOld) {
        const OpFixup op = { g, L };
        outline_cells(g, pix, bg.n_pts, lane, op);
      }
      if (!(u.layer_flags & kPoseNoInsert)) {
        __syncwarp(); // every lane is done with the old outline
        for (uint32_t i = lane; i < 2 * bg.n_pts; i += 32)
          pix[i] = fresh[i];
      }
    }
    __syncthreads();
  }
}

// The summary words of the regions a pass took bits out of (GridView::occ_dirty), recomputed from their tiles (between
// passes the words are exact or a superset, see occ_clear): a warp per word of dirty bits, lane i looks at row mask i and column mask i of the region, two ballots per layer are the two words.
__global__ void __launch_bounds__(256) k1_summaries(GridView g)
{
  const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_words = (g.n_regions + 31u) >> 5;
  for (uint32_t i = warp; i < n_words; i += n_warps) {
    uint32_t dirty = g.occ_dirty[i];
    if (!dirty)
      continue;
    __syncwarp();
    if (lane == 0)
      g.occ_dirty[i] = 0u;
    for (; dirty; dirty &= dirty - 1u) {
      const uint32_t r = i * 32u + (uint32_t)(__ffs(dirty) - 1);
      for (int L = 0; L < 2; L++) {
        const uint32_t *tile = g.occ[L] + (size_t)r * kTileWords;
        const uint32_t rows = __ballot_sync(0xffffffffu, tile[lane] != 0u), cols = __ballot_sync(0xffffffffu, tile[kRegionWidth + lane] != 0u);
        if (lane == 0)
          *reinterpret_cast<uint2 *>(g.head[r].sum[L]) = make_uint2(rows, cols);
      }
    }
  }
}

// After a block changed trees (flag raised by k1_table_updates) the owner tags are rebuilt from the cell-entry
// tables: clear, re-derive from every live entry of both layers, lower the flag - three launches in stream order,
// each of which only reads the flag when nothing changed. The host launches them after a pass that carried model
// or attribute updates, and after every gathered pass (other ranks' updates are only known on the device).
__global__ void __launch_bounds__(256) k1_owner_clear(GridView g)
{
  if (!g.head[g.n_regions].owner)
    return;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < g.n_regions; r += gridDim.x * blockDim.x)
    g.head[r].owner = kOwnerNone;
}

__global__ void __launch_bounds__(256) k1_owner_scan(GridView g)
{
  if (!g.head[g.n_regions].owner)
    return;
  const uint64_t n = (uint64_t)g.slot_mask + 1u;
  for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long cur = g.slots[i >= n][i >= n ? i - n : i];
    if (cur == kSlotEmpty || cur == kSlotTomb)
      continue;
    const uint32_t tag = (g.blk_rec[0][(uint32_t)cur - 1u].root_flags & kRecRootMask) + 1u;
    uint32_t *owner = &g.head[(uint32_t)(cur >> (32 + 2 * kRBits))].owner;
    uint32_t seen = *reinterpret_cast<volatile uint32_t *>(owner);
    if (seen == kOwnerNone)
      seen = atomicCAS(owner, kOwnerNone, tag);
    if (seen != kOwnerNone && seen != tag && seen != kOwnerMixed)
      atomicExch(owner, kOwnerMixed);
  }
}

__global__ void k1_owner_done(GridView g) { g.head[g.n_regions].owner = 0u; }

// Compaction of one layer's cell-content table: re-insert the live entries into a clean table
// (tombstones only ever turn into entries again, never into empty slots, so probe chains grow
// until this runs). Host policy: rt_api.cu.
__global__ void __launch_bounds__(256) k1_rehash(const unsigned long long *src, unsigned long long *dst, uint32_t slot_mask)
{
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= slot_mask; i += stride) {
    const unsigned long long cur = src[i];
    if (cur == kSlotEmpty || cur == kSlotTomb)
      continue;
    uint32_t h = slot_home((uint32_t)(cur >> 32), slot_mask);
    for (uint32_t probes = 0; probes <= slot_mask; probes++) { // the destination is as large as the source: a free slot exists
      if (atomicCAS(dst + h, kSlotEmpty, cur) == kSlotEmpty)
        break;
      h = (h + 1) & slot_mask;
    }
  }
}

// ---- grid digest ----------------------------------------------------------------------------
// Sum over every distinct (cell, block) entry of a 64-bit mix of (x, y, block, position of the block in the cell's
// list), per layer, plus a digest of the region counts: what two replicas - or a replica and a CPU
// world that computes the same sums from its own lists - compare instead of shipping the grid. The position is the number of
// distinct blocks of the cell that come earlier in Cell::blocks[layer] order (smaller seq).
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}

__global__ void __launch_bounds__(256) k1_grid_hash(GridView g, unsigned long long *out /* [5] */)
{
  unsigned long long h[2] = { 0ull, 0ull }, n[2] = { 0ull, 0ull }, hc = 0ull;
  const uint64_t slots_n = (uint64_t)g.slot_mask + 1u;
  for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * slots_n; i += (uint64_t)gridDim.x * blockDim.x) {
    const int L = i >= slots_n;
    const uint32_t at = (uint32_t)(L ? i - slots_n : i);
    const unsigned long long e = g.slots[L][at];
    if (e == kSlotEmpty || e == kSlotTomb)
      continue;
    const uint32_t key = (uint32_t)(e >> 32), blk = (uint32_t)e - 1u;
    const unsigned long long seq = g.blk_rec[L][blk].seq;
    // walk the cell's probe sequence: am I the first copy of this (cell, block)? how many distinct blocks precede me?
    bool first = true;
    uint32_t pos = 0;
    uint32_t hh = slot_home(key, g.slot_mask);
    for (uint32_t probes = 0; probes <= g.slot_mask; probes++, hh = (hh + 1) & g.slot_mask) {
      const unsigned long long o = g.slots[L][hh];
      if (o == kSlotEmpty)
        break;
      if ((uint32_t)(o >> 32) != key)
        continue;
      if (o == e) {
        if (hh == at)
          continue;
        // another copy of the same entry: the one met first in probe order stands for all of them
        uint32_t d_me = (at - slot_home(key, g.slot_mask)) & g.slot_mask, d_o = (hh - slot_home(key, g.slot_mask)) & g.slot_mask;
        if (d_o < d_me)
          first = false;
        continue;
      }
      const uint32_t ob = (uint32_t)o - 1u;
      const unsigned long long os = g.blk_rec[L][ob].seq;
      if (os < seq || (os == seq && ob < blk)) {
        // count a preceding block once: only at its own first copy in probe order
        bool o_first = true;
        uint32_t h2 = slot_home(key, g.slot_mask);
        for (uint32_t p2 = 0; p2 <= g.slot_mask && h2 != hh; p2++, h2 = (h2 + 1) & g.slot_mask)
          if (g.slots[L][h2] == o) {
            o_first = false;
            break;
          }
        pos += o_first ? 1u : 0u;
      }
    }
    if (!first)
      continue;
    const uint32_t region = key >> (2 * kRBits), cy = (key >> kRBits) & (kRegionWidth - 1), cx = key & (kRegionWidth - 1);
    const int32_t x = (g.rx0 + (int32_t)(region % (uint32_t)g.nrx)) * kRegionWidth + (int32_t)cx;
    const int32_t y = (g.ry0 + (int32_t)(region / (uint32_t)g.nrx)) * kRegionWidth + (int32_t)cy;
    const unsigned long long a = ((unsigned long long)(uint32_t)x << 32) | (uint32_t)y, b = ((unsigned long long)blk << 32) | pos;
    h[L] += mix64(mix64(a) ^ (b * 0x9e3779b97f4a7c15ull));
    n[L]++;
  }
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < g.n_regions; r += gridDim.x * blockDim.x) {
    const uint32_t c = g.head[r].count;
    if (c) {
      const int32_t rx = g.rx0 + (int32_t)(r % (uint32_t)g.nrx), ry = g.ry0 + (int32_t)(r / (uint32_t)g.nrx);
      hc += mix64(mix64(((unsigned long long)(uint32_t)rx << 32) | (uint32_t)ry) ^ ((unsigned long long)c * 0x9e3779b97f4a7c15ull));
    }
  }
  for (int off = 16; off; off >>= 1) {
    h[0] += __shfl_down_sync(0xffffffffu, h[0], off);
    h[1] += __shfl_down_sync(0xffffffffu, h[1], off);
    n[0] += __shfl_down_sync(0xffffffffu, n[0], off);
    n[1] += __shfl_down_sync(0xffffffffu, n[1], off);
    hc += __shfl_down_sync(0xffffffffu, hc, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 0, h[0]);
    atomicAdd(out + 1, h[1]);
    atomicAdd(out + 2, n[0]);
    atomicAdd(out + 3, n[1]);
    atomicAdd(out + 4, hc);
  }
}

static int delta_grid_blocks();

// ---- growing the extent (World::AddSuperRegion, world.cc:1072-1093: the reference's world grows whenever a polygon is
// rendered outside it) -------------------------------------------------------------------------------------------
// The grid is dense over a rectangle of SuperRegions; growing it means new arrays over the larger rectangle and every
// region's count, owner tag and occupancy tiles copied to its new index, then every cell entry re-keyed (a cell key
// starts with the region index) into fresh tables. Rare (a robot driving off the map), O(grid), entirely on the device.
__global__ void __launch_bounds__(256) k1_grow_regions(GridView from, GridView to)
{
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < from.n_regions; r += gridDim.x * blockDim.x) {
    const int32_t rx = from.rx0 + (int32_t)(r % (uint32_t)from.nrx), ry = from.ry0 + (int32_t)(r / (uint32_t)from.nrx);
    const uint32_t rn = (uint32_t)((ry - to.ry0) * to.nrx + (rx - to.rx0));
    to.head[rn] = from.head[r];
    if (from.head[r].count)
      for (int L = 0; L < 2; L++)
        for (int w = 0; w < kTileWords; w++)
          to.occ[L][(size_t)rn * kTileWords + w] = from.occ[L][(size_t)r * kTileWords + w];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    to.head[to.n_regions].owner = from.head[from.n_regions].owner; // the "owners are stale" flag
}

__global__ void __launch_bounds__(256) k1_grow_rehash(GridView from, GridView to, int L)
{
  for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= from.slot_mask; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long cur = from.slots[L][i];
    if (cur == kSlotEmpty || cur == kSlotTomb)
      continue;
    const uint32_t key = (uint32_t)(cur >> 32), r = key >> (2 * kRBits);
    const int32_t rx = from.rx0 + (int32_t)(r % (uint32_t)from.nrx), ry = from.ry0 + (int32_t)(r / (uint32_t)from.nrx);
    const uint32_t rn = (uint32_t)((ry - to.ry0) * to.nrx + (rx - to.rx0));
    const uint32_t nkey = (rn << (2 * kRBits)) | (key & ((1u << (2 * kRBits)) - 1u));
    const unsigned long long entry = ((unsigned long long)nkey << 32) | (cur & 0xffffffffull);
    uint32_t h = slot_home(nkey, to.slot_mask);
    for (uint32_t probes = 0; probes <= to.slot_mask; probes++) {
      if (atomicCAS(to.slots[L] + h, kSlotEmpty, entry) == kSlotEmpty)
        break;
      h = (h + 1) & to.slot_mask;
    }
  }
}

void launch_grow(const GridView &from, const GridView &to, void *stream)
{
  cudaStream_t s = (cudaStream_t)stream;
  k1_grow_regions<<<delta_grid_blocks(), 256, 0, s>>>(from, to);
  for (int L = 0; L < 2; L++)
    k1_grow_rehash<<<delta_grid_blocks(), 256, 0, s>>>(from, to, L);
}

static int delta_grid_blocks()
{
  static int blocks = 0;
  if (!blocks) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    blocks = sms * 8; // 8 resident 256-thread CTAs per SM, grid-stride over the steps
  }
  return blocks;
}

// phase 0: table updates, removals.  phase 1: mask fix-up, insertions. The host may compact the
// cell-entry table between the two (every tombstone of phase 0 is reclaimed before new entries land).
void launch_apply_deltas(const GridView &g, const GeomView &gv, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                         unsigned long long epoch, int phase, int what, void *stream, uint64_t *launch_counter)
{
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = delta_grid_blocks();
  if (what & kDeltaEdgeRecords) { // edge records, block / model table updates
    if (phase == 0)
      k1_phase0<<<blocks, 256, 0, s>>>(g, blobs, n_ranks, slot_bytes, epoch);
    else
      k1_phase1<<<blocks, 256, 0, s>>>(g, blobs, n_ranks, slot_bytes);
    if (launch_counter)
      *launch_counter += 1;
  }
  if ((what & kDeltaPoseRecords) && gv.geom) { // the pose path's share of the same phase
    if (phase == 0)
      k1_pose_phase0<<<blocks, 256, 0, s>>>(g, gv, blobs, n_ranks, slot_bytes, epoch);
    else
      k1_pose_phase1<<<blocks, 256, 0, s>>>(g, gv, blobs, n_ranks, slot_bytes);
    if (launch_counter)
      *launch_counter += 1;
  }
}

void launch_summaries(const GridView &g, void *stream)
{
  const int words = (int)((g.n_regions + 31u) >> 5);
  k1_summaries<<<min(delta_grid_blocks(), (words + 7) / 8), 256, 0, (cudaStream_t)stream>>>(g);
}

void launch_owner_rebuild(const GridView &g, void *stream, uint64_t *launch_counter)
{
  cudaStream_t s = (cudaStream_t)stream;
  k1_owner_clear<<<delta_grid_blocks(), 256, 0, s>>>(g);
  k1_owner_scan<<<delta_grid_blocks(), 256, 0, s>>>(g);
  k1_owner_done<<<1, 1, 0, s>>>(g);
  if (launch_counter)
    *launch_counter += 3;
}

void launch_grid_hash(const GridView &g, unsigned long long *out, void *stream)
{
  k1_grid_hash<<<delta_grid_blocks(), 256, 0, (cudaStream_t)stream>>>(g, out);
}

void launch_rehash(const unsigned long long *src, unsigned long long *dst, uint32_t slot_mask, void *stream)
{
  k1_rehash<<<delta_grid_blocks(), 256, 0, (cudaStream_t)stream>>>(src, dst, slot_mask);
}

// ------------------------------------------------------------------------------------------
// K2a: per-beam headings
// ------------------------------------------------------------------------------------------

// K2a: the heading every beam's World::Raytrace receives. A ranger's headings are a serial
// recurrence through a float (model_ranger.cc:237-241,254): fa(t) = (float)a(t), a(t+1) = (double)fa(t)
// + incr. One warp per fan cuts it into 32 chunks. A chunk's entry state is the float of the beam
// before it, fa(t0 - 1). Every lane runs its chunk from an assumed entry state; the result stands
// only when every lane's exit state fa(t1 - 1) is, bit for bit, the entry state the next lane
// assumed - then by induction from a(0) = origin.a every chunk was computed from the true state.
// Where a link does not hold, the lane after the first broken link takes the (now certain) exit
// state of its predecessor, and the lanes beyond it re-guess by translation: within a float binade
// the recurrence adds a constant per step, so a chunk's net advance D = exit - entry does not
// depend on its entry state. At least one more link becomes certain per round (at most 32 rounds,
// the serial cost); 2-5 rounds are typical, the extra ones where a fan sweeps through binade
// borders (|a| = 2, 1, 1/2, ... and the sign change).
// kKeep: the chunk's values fa(t0 .. t0 + steps) are left in `keep` (shared memory) for the coalesced write-out.
template <bool kKeep>
__device__ __forceinline__ float ranger_chunk(bool first_lane, double oa, float entry, double incr, uint32_t steps, float *keep)
{
  double a = first_lane ? oa : (double)entry + incr;
  float fa = entry; // an empty chunk hands its entry state on
  for (uint32_t i = 0; i < steps; i++) {
    fa = (float)a;
    if (kKeep)
      keep[i] = fa;
    a = (double)fa + incr;
  }
  return fa;
}

constexpr uint32_t kHeadingsKeep = 2048; // fans up to this many samples keep their chain in shared memory

// A first guess of fa(T), good enough to be right almost always (the link verification below decides): while fa
// stays inside one float binade, with ulp u, every step adds the same d = incr rounded to a multiple of u (fa is a
// multiple of u, so RN_float(fa + incr) = fa + d), and the guess jumps to the last value inside the binade in one
// go; the step that leaves the binade is taken as the recurrence takes it. A 270-degree fan passes through some
// twenty binades (|a| = 2, 1, 1/2 ... on either side of 0), so this costs twenty short iterations instead of the
// one round of the verification loop per binade border that the ideal sweep oa + T * incr used to cost.
__device__ __forceinline__ float guess_fa(double oa, double incr, uint32_t T)
{
  float f = (float)oa;
  uint32_t t = 0;
  for (int iter = 0; iter < 96 && t < T; iter++) {
    const uint32_t eb = (__float_as_uint(f) >> 23) & 0xffu;
    if (eb > 1u && eb < 0xffu && incr > 0.0) {
      const double u = __hiloint2double((int)((eb - 150u + 1023u) << 20), 0); // 2^(eb - 127 - 23)
      const double d = rint(incr / u) * u;
      if (d == 0.0)
        return f; // incr is below half an ulp: fa never moves again
      const double fd = (double)f;
      double k_in; // steps whose result is still inside this binade
      if (f > 0.0f)
        k_in = ceil((ldexp(u, 24) - fd) / d) - 1.0; // first k with fd + k d >= 2^(eb - 126)
      else
        k_in = floor((-ldexp(u, 23) - fd) / d); // last k with fd + k d <= -2^(eb - 127)
      if (k_in >= 1.0) {
        const uint32_t k = (uint32_t)fmin(k_in, (double)(T - t));
        f = (float)(fd + (double)k * d);
        t += k;
      }
    }
    if (t < T) { // the step that leaves the binade (or any step near zero), as the recurrence takes it
      f = (float)((double)f + incr);
      t++;
    }
  }
  return f;
}

__global__ void __launch_bounds__(32) k2_fan_headings(FanBatchView b)
{
  const uint32_t f = blockIdx.x, lane = threadIdx.x;
  const FanRec fan = b.fans[f];
  const uint32_t first = b.fan_first_beam[f], n = fan.n_samples;
  double *heading = b.heading + first;
  double *bearing = b.bearings ? b.bearings + first : nullptr;
  __shared__ float chain[kHeadingsKeep];
  if (fan.angles == 0 /* STG_RT_ANGLES_RANGER */) {
    const uint32_t nm1 = n - 1u; // unsigned, like std::max(sample_count - 1, 1u)
    const double incr = fan.fov / (double)(nm1 > 1u ? nm1 : 1u);
    const double start = n > 1 ? -fan.fov / 2.0 : 0.0;
    const uint32_t chunk = (n + 31u) / 32u;
    const bool keep = n <= kHeadingsKeep;
    const uint32_t t0 = min(lane * chunk, n), t1 = min(t0 + chunk, n); // this lane's beams
    // entry state fa(t0 - 1): first guess binade by binade (lane 0 starts from origin.a itself)
    float entry = t0 > 0 ? guess_fa(fan.oa, incr, t0 - 1u) : 0.0f;
    for (int round = 0; round < 33; round++) {
      const float exit = keep ? ranger_chunk<true>(lane == 0, fan.oa, entry, incr, t1 - t0, chain + t0)
                              : ranger_chunk<false>(lane == 0, fan.oa, entry, incr, t1 - t0, nullptr);
      const float next_entry = __shfl_down_sync(0xffffffffu, entry, 1);
      const bool link_ok = lane == 31 || __float_as_uint(exit) == __float_as_uint(next_entry);
      const uint32_t broken = __ballot_sync(0xffffffffu, !link_ok);
      if (!broken)
        break;
      const uint32_t j = __ffs(broken) - 1; // lanes 0..j hold the true state; so does the exit of lane j
      // inclusive scan of the advances of lanes j+1.. (floats and their differences: exact in double)
      double adv = lane > j ? (double)exit - (double)entry : 0.0;
      for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, adv, d);
        if (lane >= (uint32_t)d)
          adv += up;
      }
      const float anchor = __shfl_sync(0xffffffffu, exit, j);    // true entry state of lane j+1
      const double before = __shfl_up_sync(0xffffffffu, adv, 1); // sum of the advances of lanes j+1..lane-1
      if (lane == j + 1)
        entry = anchor;
      else if (lane > j + 1)
        entry = (float)((double)anchor + before);
    }
    // emit (model_ranger.cc:237-238): float savedAngle; float distortedAngle = a + incr * angle_noise *
    // simpleNoise() * 0.5 - the noise term is exactly 0 without jitter, as there
    const double jitter = b.noise ? b.noise[f].angle_noise : 0.0;
    const unsigned long long seed = b.noise ? (b.noise[f].seed ? b.noise[f].seed : fan.finder + 1ull) : 0ull;
    if (keep) { // the verified chain is in shared memory: write it out a warp-wide line at a time
      __syncwarp();
      for (uint32_t t = lane; t < n; t += 32) {
        if (jitter != 0.0) {
          const double a = t == 0 ? fan.oa : (double)chain[t - 1] + incr;
          heading[t] = (double)(float)(a + incr * jitter * noise_simple(noise_bits(seed, first + t, b.noise_epoch, 0)) * 0.5);
        } else {
          heading[t] = (double)chain[t];
        }
      }
    } else {
      double a = lane == 0 ? fan.oa : (double)entry + incr;
      for (uint32_t t = t0; t < t1; t++) {
        const float fa = (float)a;
        heading[t] = jitter != 0.0 ? (double)(float)(a + incr * jitter * noise_simple(noise_bits(seed, first + t, b.noise_epoch, 0)) * 0.5)
                                   : (double)fa;
        a = (double)fa + incr;
      }
    }
    if (bearing)
      for (uint32_t t = lane; t < n; t += 32)
        bearing[t] = start + ((double)t) * incr;
  } else if (fan.angles == 1 /* STG_RT_ANGLES_EVEN_FOV */) {
    const double starta = fan.fov / 2.0 - fan.oa;
    for (uint32_t s = lane; s < n; s += 32) {
      const double h = ((double)s * fan.fov / (double)(n - 1u)) - starta;
      heading[s] = h;
      if (bearing)
        bearing[s] = h - fan.oa;
    }
  } else {
    for (uint32_t s = lane; s < n; s += 32) {
      const double h = b.headings_in[fan.first_heading + s];
      heading[s] = h;
      if (bearing)
        bearing[s] = h - fan.oa;
    }
  }
}

// The same headings for batches of SHORT fans (single-beam sonars, the one-ray fans of the fiducial and
// explicit-ray paths: BASELINE config 4 has 1.6 M of them): one thread per fan, the recurrence run as it is.
__global__ void __launch_bounds__(128) k2_fan_headings_short(FanBatchView b)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= b.n_fans)
    return;
  const FanRec fan = b.fans[f];
  const uint32_t first = b.fan_first_beam[f], n = fan.n_samples;
  double *heading = b.heading + first;
  double *bearing = b.bearings ? b.bearings + first : nullptr;
  if (fan.angles == 0 /* STG_RT_ANGLES_RANGER */) {
    const uint32_t nm1 = n - 1u;
    const double incr = fan.fov / (double)(nm1 > 1u ? nm1 : 1u);
    const double start = n > 1 ? -fan.fov / 2.0 : 0.0;
    const double jitter = b.noise ? b.noise[f].angle_noise : 0.0;
    const unsigned long long seed = b.noise ? (b.noise[f].seed ? b.noise[f].seed : fan.finder + 1ull) : 0ull;
    double a = fan.oa;
    for (uint32_t t = 0; t < n; t++) {
      const float fa = (float)a;
      heading[t] = jitter != 0.0 ? (double)(float)(a + incr * jitter * noise_simple(noise_bits(seed, first + t, b.noise_epoch, 0)) * 0.5)
                                 : (double)fa;
      a = (double)fa + incr;
      if (bearing)
        bearing[t] = start + ((double)t) * incr;
    }
  } else if (fan.angles == 1 /* STG_RT_ANGLES_EVEN_FOV */) {
    const double starta = fan.fov / 2.0 - fan.oa;
    for (uint32_t s = 0; s < n; s++) {
      const double h = ((double)s * fan.fov / (double)(n - 1u)) - starta;
      heading[s] = h;
      if (bearing)
        bearing[s] = h - fan.oa;
    }
  } else {
    for (uint32_t s = 0; s < n; s++) {
      const double h = b.headings_in[fan.first_heading + s];
      heading[s] = h;
      if (bearing)
        bearing[s] = h - fan.oa;
    }
  }
}

// ModelRanger::Sensor::Update's set-up (model_ranger.cc:217-230) for every sensor of every ranger model: the fan
// record stg_rt_fans_submit would have been given. Pose::operator+ (stage.hh:297-304) with the host C library's
// cos / sin (rt_libm.cuh), same operation order, no contraction.
__global__ void __launch_bounds__(128) k2_expand_rangers(const RangerSensorRec *sensors, uint32_t n_sensors, const uint32_t *sensor_first,
                                                         const RangerPoseRec *poses, uint32_t n_rangers, FanRec *fans, uint32_t *first)
{
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_fans = (uint64_t)n_rangers * n_sensors;
  const uint32_t per_model = sensor_first[n_sensors];
  if (idx == 0)
    first[n_fans] = (uint32_t)((uint64_t)n_rangers * per_model);
  if (idx >= n_fans)
    return;
  const uint32_t r = (uint32_t)(idx / n_sensors), s = (uint32_t)(idx % n_sensors);
  const RangerSensorRec sn = sensors[s];
  const RangerPoseRec P = poses[r];
  const double start_angle = sn.n_samples > 1 ? -sn.fov / 2.0 : 0.0; // :221
  const double la = sn.pa + start_angle, lz = sn.pz + sn.size_z / 2.0; // :224-225
  double cosa, sina;
  if (libm::covered(P.a)) {
    sina = libm::sin_host_exact(P.a);
    cosa = libm::cos_host_exact(P.a);
  } else {
    sina = sin(P.a);
    cosa = cos(P.a);
  }
  double a = P.a + la; // normalize(), stage.hh:163-170
  while (a < -3.14159265358979323846)
    a += 2.0 * 3.14159265358979323846;
  while (a > 3.14159265358979323846)
    a -= 2.0 * 3.14159265358979323846;
  FanRec f;
  f.finder = P.finder;
  f.n_samples = sn.n_samples;
  f.pred = 0; // STG_RT_PRED_RANGER
  f.ztest = 1;
  f.layer = P.layer;
  f.angles = 0; // STG_RT_ANGLES_RANGER
  f.first_heading = 0;
  f.ox = P.x + sn.px * cosa - sn.py * sina;
  f.oy = P.y + sn.px * sina + sn.py * cosa;
  f.oz = P.z + lz;
  f.oa = a;
  f.range = sn.range_max;
  f.fov = sn.fov;
  fans[idx] = f;
  first[idx] = r * per_model + sensor_first[s];
}

void launch_expand_rangers(const RangerSensorRec *sensors, uint32_t n_sensors, const uint32_t *sensor_first, const RangerPoseRec *poses,
                           uint32_t n_rangers, FanRec *fans, uint32_t *first, void *stream)
{
  const uint64_t n = (uint64_t)n_rangers * n_sensors;
  if (n)
    k2_expand_rangers<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(sensors, n_sensors, sensor_first, poses, n_rangers, fans, first);
}

void launch_fan_headings(const FanBatchView &b, void *stream)
{
  if (!b.n_fans)
    return;
  if (b.n_beams <= 8ull * b.n_fans) // short fans on average: a warp per fan would idle
    k2_fan_headings_short<<<(b.n_fans + 127) / 128, 128, 0, (cudaStream_t)stream>>>(b);
  else
    k2_fan_headings<<<b.n_fans, 32, 0, (cudaStream_t)stream>>>(b);
}

// Diagnostics: the store pattern the march uses for its results - every warp writes 256 contiguous bytes - aimed at
// page-locked host memory, to measure what the link to the host takes (stg_rt_debug_link_bandwidth).
__global__ void __launch_bounds__(128) k_host_store(unsigned long long *dst, size_t n_words, unsigned long long seed)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = seed + i;
}

void launch_host_store(void *dst, size_t n_words, unsigned long long seed, void *stream)
{
  k_host_store<<<148 * 16, 128, 0, (cudaStream_t)stream>>>((unsigned long long *)dst, n_words, seed);
}

// K2, the ray march, lives in rt_march.cu.

} // namespace stg
