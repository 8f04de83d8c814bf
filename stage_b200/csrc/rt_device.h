// Device-side data layout shared by the kernels (rt_kernels.cu) and the host runtime (rt_api.cu).
// See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

namespace stg {

constexpr int kRBits = 5;                    // region.hh:13 RBITS: a Region is 32 x 32 cells
constexpr int kRegionWidth = 1 << kRBits;    // region.hh:17
constexpr int kTileWords = 2 * kRegionWidth; // row masks then column masks
// words of the GridView::occ allocation for n regions: both layers' tiles, then the dirty bits
constexpr size_t occ_alloc_words(size_t n) { return 2 * n * kTileWords + (n + 31) / 32; }
constexpr int kSRBits = 10;                  // region.hh:15 SRBITS: a SuperRegion is 1024 x 1024 cells

constexpr unsigned long long kSlotEmpty = 0ull;
constexpr unsigned long long kSlotTomb = ~0ull;
constexpr uint32_t kSectorSlots = 4; // slots per 32-byte sector

#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t mix32(uint32_t k)
{
  k ^= k >> 16;
  k *= 0x7feb352du;
  k ^= k >> 15;
  k *= 0x846ca68bu;
  k ^= k >> 16;
  return k;
}

// Where the probe sequence of a cell key starts: at a 32-byte sector (4 slots), which the lookup
// reads whole; from there on it is linear probing, slot by slot. (Tried and dropped: one 64-byte
// bucket per 4 x 4-cell patch, so that neighbouring beams' lookups share sectors - fewer DRAM
// sectors but longer probe runs; the march and K1 both got slower.)
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t slot_home(uint32_t key, uint32_t slot_mask)
{
  return (mix32(key) * kSectorSlots) & slot_mask;
}

// The replicated world state one GPU holds. Passed to kernels by value.
//
//  (reg_count, reg_owner and occ_sum are fields of head[r], see RegionHead)
//  reg_count[r]        Region::count (region.cc:19-34): number of (cell, block) entries of BOTH
//                      layers in region r. r = (gry - ry0) * nrx + (grx - rx0), row-major over the
//                      dense extent. 4 B / region.
//  reg_owner[r]        whose entries region r holds, as far as that is one tree: 0 = nothing, root + 1 = every
//                      entry (both layers) added since the region was last empty belongs to a model of
//                      the tree with that root, kOwnerMixed = more than one tree (or unknown). A ray
//                      whose finder belongs to that tree cannot be stopped there by a predicate that
//                      rejects related models (Model::IsRelated, model.cc:446-466), so the march walks
//                      such a region without looking at its masks or cells: a robot alone in its
//                      regions does not look up its own outline. Conservative by construction: any
//                      doubt is kOwnerMixed, which is the ordinary walk. 4 B / region.
//  occ[L][r*64 + ..]   occupancy tile of region r, layer L, 256 bytes = two 128-byte lines:
//                      words 0..31  row masks:    bit cx of word cy      is set iff the cell's block
//                      words 32..63 column masks: bit cy of word 32 + cx   list Cell::blocks[L]
//                      (region.hh:47) is non-empty. This is what the ray march reads: an x-major
//                      ray tests a whole run of cells of one row with one AND, a y-major ray does
//                      the same on a column.
//  occ_sum[L][r*2 ..]  two summary words per tile: bit cy of word 0 is set iff row mask cy is non-zero, bit cx of
//                      word 1 iff column mask cx is (the tile's projections on the two axes). The march uses them to
//                      leave a region whose occupied cells lie off its path in closed form, and to jump to the first
//                      row / column that can hold a hit (rt_walk.h). Kept in the region's RegionHead.
//                      Between K1 passes a word may hold bits too many (never too few): see occ_dirty.
//  occ_dirty[r / 32]   bit r % 32: a K1 pass cleared a bit of region r's tiles and left the summary bits standing;
//                      k1_summaries recomputes those regions' words after the pass, beside the march. Same allocation as the tiles, behind them.
//  slots[L][h]         open-addressing multimap (cell -> block) holding the CONTENT of the occupied
//                      cells: entry = (cell_key << 32) | (block_id + 1), cell_key = r*1024 + cy*32
//                      + cx. 0 = empty, ~0 = tombstone. Only consulted where an occ bit is set.
//                      Probing starts at slot_home(cell_key) (above).
//  blk_rec[L][b]       BlockRec below: what the hit test reads about block b (Block::global_z
//                      block.cc:177-180, owning model, tree root for Model::IsRelated, Model::vis bits
//                      stage.hh:1924-1937) plus seq, the order key of b's latest Map on layer L:
//                      entries of one cell sorted by it reproduce the order of Cell::blocks[L].
//  mdl_*               per model: root (for the finder), ranger_return (intensity output), vis flags.
// Everything the hit test needs about one block on one layer, in ONE 32-byte sector, so resolving an
// occupied cell costs two dependent loads (slot, then this) instead of a chain through block and
// model tables. zmin/zmax/model/root_flags are per block (kept equal on both layers); seq is per layer.
struct BlockRec {
  double zmin, zmax;       // Block::global_z (block.cc:177-180)
  unsigned long long seq;  // order key of the latest Map on this layer
  uint32_t model;          // owning model id
  uint32_t root_flags;     // root id of the model's tree | kRecObstacle | kRecNoRanger | kRecGripper
};
static_assert(sizeof(BlockRec) == 32, "BlockRec");
constexpr uint32_t kRecNoRanger = 1u << 30; // vis.ranger_return < 0: invisible to rangers
constexpr uint32_t kRecGripper = 1u << 31;  // vis.gripper_return
constexpr uint32_t kRecObstacle = 1u << 29; // vis.obstacle_return
constexpr uint32_t kRecRootMask = (1u << 29) - 1u;
constexpr uint32_t kOwnerNone = 0u, kOwnerMixed = 0xffffffffu;
// BlockRec::seq = epoch (36 bits) | rank (8 bits) | call order within the rank's blob (20 bits): the order of the
// latest Map across the whole job. An epoch is one K1 pass; 2^36 of them is 79 days at 10,000 passes a second.
constexpr int kSeqEpochShift = 28, kSeqRankShift = 20;
constexpr uint32_t kSeqLocalMax = (1u << kSeqRankShift) - 1u;
// device-side error bits (GridView::err_host, a word of page-locked host memory the kernels OR into)
constexpr uint32_t kErrTableFull = 1u, kErrBadBlob = 2u, kErrOutOfExtent = 4u;
constexpr uint32_t kOwnerNever = 0xfffffffeu; // a finder tag that matches no region

// What the march asks about a region before it looks at its tile, in ONE 32-byte sector: the probe of a region
// (Region::count) brings the owner tag and the tile summaries of both layers along, so the two or three questions a
// populated region answers cost one trip to L2 or DRAM, and K1's updates of them land in one sector as well.
struct RegionHead {
  uint32_t count;     // reg_count
  uint32_t owner;     // reg_owner
  uint32_t sum[2][2]; // occ_sum: per layer, which row masks / which column masks of the tile are non-zero
  uint32_t pad[2];
};
static_assert(sizeof(RegionHead) == 32, "RegionHead");

struct GridView {
  int32_t rx0, ry0, nrx, nry;
  double ppm;
  RegionHead *head;      // n_regions records, then one more whose `owner` is the "owners are stale" flag (launch_owner_rebuild)
  uint32_t n_regions;
  uint32_t *occ[2];
  uint32_t *occ_dirty;
  unsigned long long *slots[2];
  uint32_t slot_mask;
  uint32_t n_blocks, n_models;
  BlockRec *blk_rec[2];
  uint32_t *mdl_root;
  double *mdl_ranger_return;
  uint32_t *mdl_flags;
  uint8_t *mdl_pal;      // per model: index of its ranger_return in the host's palette of distinct values (0 is "no hit")
  uint32_t *err_host;    // page-locked host word: kErr* bits raised by K1 (read by stg_rt_sync)
};

// ---- moved-block delta blob (host -> device, and rank -> rank through the all-gather) ----------
// [DeltaHeader][EdgeRec x (n_remove + n_insert)][BlockUpd x n_block_upd][ModelUpd x n_model_upd][PoseUpd x n_pose_upd]
struct DeltaHeader {
  uint32_t n_remove, n_insert;       // edge records: removals first, then insertions
  uint32_t n_block_upd, n_model_upd;
  uint32_t steps_remove, steps_insert; // total cell visits of the remove / insert edges
  uint32_t rank, magic;              // kDeltaMagic: K1 refuses a slot that does not carry it (a torn or stale gather)
  uint32_t steps_remove_l[2], steps_insert_l[2]; // the same per layer: every replica books the other ranks' net change
  uint32_t n_bytes;                  // size of the whole blob; K1 checks the counts against it and against the slot
  uint32_t n_pose_upd;               // PoseUpd records after the model updates (outlines derived on the device)
  unsigned long long epoch;          // the sender's next epoch; a gathered pass uses the largest one it finds
  // cell visits of the outlines the pose records take out / put in, per layer, as far as the sender can tell without
  // computing them (an upper estimate from the blocks' perimeters): for the receivers' table accounting only
  uint32_t est_remove_l[2], est_insert_l[2];
  uint32_t reserved[4];
};
static_assert(sizeof(DeltaHeader) == 96, "DeltaHeader");
constexpr uint32_t kDeltaMagic = 0x53544744u; // "STGD"

// Block::UnMap(layer) + Block::Map(layer) of one block whose outline the DEVICE derives from the model's pose
// (Model::LocalToPixels, model.cc:474-491): x, y, z, a = GetGlobalPose() + geom.pose. 48 bytes instead of the 8 edge
// records + 1 block update (296 bytes) the same move costs with host-rasterised outlines.
struct PoseUpd {
  uint32_t block;
  uint32_t layer_flags; // layer | kPose*
  uint32_t seq_local;   // order of the Map call within this rank's blob
  uint32_t root_flags;  // BlockRec::root_flags of the block's model, as the sender knows it
  double x, y, z, a;
};
static_assert(sizeof(PoseUpd) == 48, "PoseUpd");
constexpr uint32_t kPoseRemoveOld = 0x100u; // the device holds this block's outline on that layer (GeomView::pix): take it out first
constexpr uint32_t kPoseNoInsert = 0x200u;  // UnMap only

// Local geometry of a block, registered once (stg_rt_block_define): its polygon in model coordinates as Block::pts
// holds it after BlockGroup::CalcSize (blockgroup.cc:84-112) and Block::local_z.
struct BlockGeom {
  uint32_t first_pt, n_pts; // into GeomView::pts / pix (2 values per point)
  uint32_t model, est_steps; // est_steps: upper estimate of the outline's cell visits at any pose (how many warps its rounds of 32 can use)
  double zmin, zmax;
};
static_assert(sizeof(BlockGeom) == 32, "BlockGeom");

struct GeomView {
  const BlockGeom *geom;  // [n_blocks]
  const double *pts;      // local points, x then y
  int32_t *pix[2];        // per layer: the pixel outline the device rendered last for each defined block (valid while it is mapped through a PoseUpd)
  int32_t *pix_new[2];    // per layer: outlines computed in phase 0 of the running pass, committed to pix in phase 1
  uint32_t n_blocks;      // blocks the table covers
};

// One polygon edge = one integer line of World::MapPoly (world.cc:995-1052): start cell included,
// end cell excluded, n_steps = |dx| + |dy| cell visits. first_step = running sum within its group.
struct EdgeRec {
  int32_t x0, y0, x1, y1;
  uint32_t block;
  uint32_t layer;
  uint32_t first_step, n_steps;
};
static_assert(sizeof(EdgeRec) == 32, "EdgeRec");

struct BlockUpd {
  uint32_t block, model;
  double zmin, zmax;
  uint32_t seq_local;  // order of the Map call within this rank's blob
  uint32_t layer;      // 0/1, | kUpdKeepSeq when only the attributes changed (no new Map)
  uint32_t root_flags; // BlockRec::root_flags
  uint32_t pad;
};
static_assert(sizeof(BlockUpd) == 40, "BlockUpd");
constexpr uint32_t kUpdKeepSeq = 0x100u;

struct ModelUpd {
  uint32_t id, root;
  double ranger_return;
  int32_t fiducial_return, fiducial_key;
  uint32_t flags, pal; // pal: palette index of ranger_return (GridView::mdl_pal)
};
static_assert(sizeof(ModelUpd) == 32, "ModelUpd");

// ---- fans -----------------------------------------------------------------------------------
// Device copy of stg_rt_fan (include/stg_rt.h) - same 64-byte layout.
struct FanRec {
  uint32_t finder, n_samples;
  uint8_t pred, ztest, layer, angles;
  uint32_t first_heading;
  double ox, oy, oz, oa, range, fov;
};
static_assert(sizeof(FanRec) == 64, "FanRec");

// stg_rt_fan_noise
struct FanNoiseRec {
  double range_noise_const, range_noise, angle_noise;
  unsigned long long seed;
};
static_assert(sizeof(FanNoiseRec) == 32, "FanNoiseRec");

// Counter-based random numbers for the sensor noise: a 64-bit mix of (seed, beam, launch, draw).
__host__ __device__ inline unsigned long long noise_bits(unsigned long long seed, unsigned long long beam, unsigned long long epoch,
                                                         unsigned draw)
{
  unsigned long long z = seed * 0x9e3779b97f4a7c15ull + beam * 0xbf58476d1ce4e5b9ull + epoch * 0x94d049bb133111ebull + draw;
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}
// simpleNoise(), model_ranger.cc:173-177: 2 * (rand() % 1000 * 0.001 - 0.5)
__host__ __device__ inline double noise_simple(unsigned long long bits) { return 2 * ((double)(bits % 1000ull) * 0.001 - 0.5); }

struct FanBatchView {
  uint32_t n_fans;
  uint64_t n_beams;
  uint64_t beam_begin, beam_end; // the slice of beams one march launch covers (chunked launches)
  const FanRec *fans;
  const uint32_t *fan_first_beam; // n_fans + 1 prefix sums
  uint32_t uniform_samples;       // n_samples if every fan has the same, else 0 (beam -> fan by division)
  const double *headings_in;      // EXPLICIT headings (may be null)
  const double *trig_in;          // strict mode (cos, sin) per beam (may be null)
  const FanNoiseRec *noise;       // per fan sensor noise (may be null)
  uint64_t noise_epoch;           // changes every launch, so every tick draws fresh noise
  double *heading;                // per beam: the heading World::Raytrace receives
  // outputs (any may be null)
  double *ranges, *intensities, *bearings;
  int32_t *hit_model, *hit_cell;
  uint32_t *steps;
  uint8_t *intens_idx; // intensities as one palette index per beam (GridView::mdl_pal): 1 byte over PCIe instead of 8
  // Streamed delivery. The batch is cut into stretches of 2^stretch_shift beams; stretch_count holds one DEVICE word per
  // stretch, stretch_done one word of PAGE-LOCKED HOST memory. A warp of the march, once its 32 results have been stored
  // (into page-locked host memory) and fenced system-wide, adds one to its stretch's counter; the warp that completes a
  // stretch puts the counter back to zero and writes done_seq into the stretch's host word, so that host threads can start
  // on the stretch while the kernel is still marching the rest - one small write over the link per stretch, not per warp.
  // Null: nobody is listening.
  uint32_t *stretch_count;
  uint32_t *stretch_done;
  uint32_t done_seq, stretch_shift;
};

// ---- whole ranger models (same layouts as stg_rt_ranger_* in include/stg_rt.h) ------------------
struct RangerSensorRec {
  double px, py, pz, pa;
  double size_z, range_max, fov;
  uint32_t n_samples, reserved;
};
static_assert(sizeof(RangerSensorRec) == 64, "RangerSensorRec");
struct RangerPoseRec {
  uint32_t finder;
  uint8_t layer, pad[3];
  double x, y, z, a;
};
static_assert(sizeof(RangerPoseRec) == 40, "RangerPoseRec");
// writes n_rangers * n_sensors fan records and their first-beam prefix (n_fans + 1 entries)
void launch_expand_rangers(const RangerSensorRec *sensors, uint32_t n_sensors, const uint32_t *sensor_first, const RangerPoseRec *poses,
                           uint32_t n_rangers, FanRec *fans, uint32_t *first, void *stream);

// ---- collision tests (Model::TestCollision) ----------------------------------------------------
// One query = one top-level Model::TestCollision: the outlines (EdgeRec, first_step running within
// the query) of the model's blocks and its descendants' in visiting order.
struct CollisionQuery {
  uint32_t first_edge, n_edges;
  uint32_t n_steps; // cell visits of all its edges
  uint32_t layer;
};
static_assert(sizeof(CollisionQuery) == 16, "CollisionQuery");

struct CollisionBatchView {
  uint32_t n_queries;
  const CollisionQuery *queries;
  const EdgeRec *edges;
  int32_t *hit_model; // per query: id of the first model in the way, -1 for none
};
void launch_collisions(const GridView &g, const CollisionBatchView &b, void *stream);

// Model::AppendTouchingModels: same queries (own blocks only), a set of models per query
constexpr uint32_t kTouchersCap = 120; // distinct models one query can report
struct ToucherBatchView {
  uint32_t n_queries, max_per_query;
  const CollisionQuery *queries;
  const EdgeRec *edges;
  uint32_t *out;       // [n_queries * max_per_query] model ids, ascending
  uint32_t *out_count; // per query: how many there are; 0xffffffff: more than kTouchersCap
};
void launch_touchers(const GridView &g, const ToucherBatchView &b, void *stream);

// ---- batched Move (ModelPosition::Move) ----------------------------------------------------------
// What a mover's new outline touches (Block::TestCollision's condition, block.cc:153-157), split into
// "something that does not move this tick" and the list of other movers it touches.
constexpr uint32_t kTouchMax = 30;
struct TouchRec {
  uint32_t static_hit;        // 1: touches a model that is not one of this call's movers
  uint32_t n_touched;         // entries of touched[] (may hold duplicates; > kTouchMax: overflowed)
  uint32_t touched[kTouchMax]; // indices of other movers touched
};
static_assert(sizeof(TouchRec) == 128, "TouchRec");

struct TouchBatchView {
  uint32_t n_queries;
  const CollisionQuery *queries; // outlines to test (need not be rendered), first_step running per query
  const EdgeRec *edges;          // EdgeRec::block = the block the outline belongs to (its model, root)
  const double *edge_z;          // per edge: zmin, zmax the block will have (Block::Map sets global_z before the test)
  const uint32_t *mover_of_model; // per model: index + 1 of the mover whose tree it belongs to, 0 = none
  TouchRec *out;
};
void launch_touches(const GridView &g, const TouchBatchView &b, void *stream);
void launch_mark_movers(const uint32_t *pairs, uint32_t n, uint32_t *table, int clear, void *stream);

// ---- fiducial finders (same layouts as stg_rt_fid_* in include/stg_rt.h) -----------------------
struct FidTargetRec {
  uint32_t model;
  int32_t fiducial_return, fiducial_key;
  uint32_t reserved;
  double x, y, z, a;
  double sx, sy, sz;
};
static_assert(sizeof(FidTargetRec) == 72, "FidTargetRec");

struct FidFinderRec {
  uint32_t finder;
  int32_t fiducial_key;
  uint8_t ignore_zloc, layer, pad[6];
  double x, y, z, a;
  double ox, oy, oz, oa;
  double max_range_anon, max_range_id, fov;
};
static_assert(sizeof(FidFinderRec) == 104, "FidFinderRec");

struct FiducialRec {
  uint32_t target;
  int32_t id;
  double range, bearing;
  double geom[4], pose[4];
};
static_assert(sizeof(FiducialRec) == 88, "FiducialRec");

struct FidBatchView {
  uint32_t n_targets, n_finders, max_per_finder;
  const FidTargetRec *targets;
  const FidFinderRec *finders;
  const FanRec *fans; // one per finder: the line-of-sight ray's constants
  FiducialRec *out;
  uint32_t *out_count;
};

// uniform grid over the fiducial targets (candidate query of the fiducial finders)
struct FidGridView {
  double x0, y0, cell;   // bin (bx, by) covers [x0 + bx*cell, x0 + (bx+1)*cell) x ...
  int32_t nbx, nby;
  uint32_t *bin_of;      // [n_targets]
  uint32_t *start;       // [n_bins + 1] counts, then exclusive prefix sums
  uint32_t *cursor;      // [n_bins]
  uint32_t *sorted;      // [n_targets] target indices grouped by bin
};

// ---- blobfinders ------------------------------------------------------------------------------
struct BlobFinderRec {
  uint32_t finder, scan_width, scan_height;
  uint8_t layer, pad[3];
  uint32_t first_color, n_colors;
  double ox, oy, oz, oa;
  double range, fov;
};
static_assert(sizeof(BlobFinderRec) == 72, "BlobFinderRec");

struct BlobRec {
  uint32_t model, left, top, right, bottom, pad;
  double range;
};
static_assert(sizeof(BlobRec) == 32, "BlobRec");

struct BlobBatchView {
  uint32_t n_finders, max_per_finder;
  const BlobFinderRec *finders;
  const uint32_t *fan_first_beam;
  const int32_t *hit_model; // the traced scan lines
  const double *ranges;
  const double *mdl_color;  // 4 doubles (rgba) per model
  const double *colors;     // 3 doubles per tracked colour
  BlobRec *out;
  uint32_t *out_count;
};

void launch_fiducials(const GridView &g, const FidBatchView &b, void *stream);
void launch_fiducials_grid(const GridView &g, const FidBatchView &b, const FidGridView &v, void *stream);
void launch_fid_fans(const FidFinderRec *finders, FanRec *fans, uint32_t n, void *stream);
void launch_fid_pack(const FidBatchView &b, uint32_t *offsets, void *packed_host, uint32_t *count_host, void *stream);
void launch_blobs(const GridView &g, const BlobBatchView &b, void *stream);
void launch_debug_trig(uint64_t n, const double *x, double *cs, void *stream);

// launch wrappers implemented in rt_kernels.cu; all asynchronous on `stream`
// what: which kinds of record the blobs may hold (a local pass knows; a gathered pass says both)
constexpr int kDeltaEdgeRecords = 1, kDeltaPoseRecords = 2;
void launch_apply_deltas(const GridView &g, const GeomView &gv, const unsigned char *blobs, int n_ranks, size_t slot_bytes,
                         unsigned long long epoch, int phase, int what, void *stream, uint64_t *launch_counter);
void launch_owner_rebuild(const GridView &g, void *stream, uint64_t *launch_counter);
void launch_summaries(const GridView &g, void *stream); // tighten GridView::occ_sum over the regions marked in occ_dirty
void launch_fan_headings(const FanBatchView &b, void *stream);
void launch_ray_march(const GridView &g, const FanBatchView &b, void *stream);
// the march with lane refill (rt_march.cu): per-beam (cos, sin) first, then the kernel
bool ray_march_refill_wanted(const FanBatchView &b);
void launch_beam_trig(const FanBatchView &b, double *trig, void *stream);
void launch_ray_march_refill(const GridView &g, const FanBatchView &b, const double *trig, uint32_t *work_counter, void *stream);

} // namespace stg
