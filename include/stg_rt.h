/* stg_rt.h - C ABI of libstg_rt.so: Stage's per-tick sensor raycast path on one B200.
 *
 * This is the boundary a lightly patched libstage binds (INTEGRATION.md shows the patch): plain C,
 * POD arguments, integer return codes, no STL / torch / CUDA types. Every entry point names the
 * reference interface it stands in for (file:line under rtv/Stage 4.3.0, libstage/).
 *
 * Model of use, mirroring one World::Update (world.cc:605-676):
 *   load time    stg_rt_create; stg_rt_model_upsert per Model; stg_rt_block_map per Block::Map
 *   every tick   sensor Update()s call stg_rt_fans_submit instead of looping World::Raytrace;
 *                Block::Map / Block::UnMap (from ModelPosition::Move, Model::SetPose ...) also call
 *                stg_rt_block_map / stg_rt_block_unmap;
 *                World::Update calls stg_rt_flush + stg_rt_sync before CallUpdateCallbacks
 *                (world.cc:668), the first point anybody reads sensor data.
 * Calls are applied in issue order: a fan sees exactly the map/unmap calls issued before it.
 *
 * Identity: models and blocks are named by dense uint32 ids chosen by the caller (Model::id,
 * stage.hh:1794, for models). Pointers never cross the boundary.
 *
 * Threading: every entry point may be called from any thread (libstage runs fiducial Update()s on
 * worker threads, world.cc:228-262); calls are serialised inside.
 *
 * Errors: 0 = ok, negative = error (STG_RT_E*); stg_rt_last_error gives text. The library never
 * calls exit() and has NO CPU fallback: if the GPU path cannot run, the call fails.
 */
#ifndef STG_RT_H
#define STG_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STG_RT_ABI_VERSION 1

enum {
  STG_RT_OK = 0,
  STG_RT_EINVAL = -1,   /* bad argument */
  STG_RT_ECUDA = -2,    /* CUDA runtime error (text in stg_rt_last_error) */
  STG_RT_ENOMEM = -3,   /* a fixed-capacity table is full (models, blocks, cell entries) */
  STG_RT_EEXTENT = -4,  /* the world would span more than 4095 SuperRegions (the device grid's limit) */
  STG_RT_ESTATE = -5,   /* call not valid in the current state */
  STG_RT_ENODEV = -6    /* no usable CUDA device */
};

/* Ray filter predicates: the reference's five static ray_test_func_t (stage.hh:569). */
enum {
  STG_RT_PRED_RANGER = 0,    /* ranger_match, model_ranger.cc:158-170: hit is outside the finder's
                                model tree and hit.vis.ranger_return >= 0 */
  STG_RT_PRED_UNRELATED = 1, /* fiducial_raytrace_match model_fiducial.cc:103-107, blob_match
                                model_blobfinder.cc:102-107, bumper_match model_bumper.cc:157-162:
                                hit is outside the finder's model tree (Model::IsRelated,
                                model.cc:446-466) */
  STG_RT_PRED_GRIPPER = 2    /* gripper_raytrace_match, model_gripper.cc:329-336:
                                hit != finder and hit.vis.gripper_return */
};

/* How the per-beam headings of a fan are generated. */
enum {
  STG_RT_ANGLES_RANGER = 0,  /* ModelRanger::Sensor::Update, model_ranger.cc:232-255: heading of
                                beam t is (double)(float)a_t, a_0 = origin.a,
                                a_{t+1} = (double)(float)a_t + fov/max(n-1,1). origin is the pose
                                of the FIRST ray (rayorg after LocalToGlobal, :226-230) */
  STG_RT_ANGLES_EVEN_FOV = 1,/* World::Raytrace(gpose,range,fov,...), world.cc:721-744: heading of
                                beam s is s*fov/(n-1) - (fov/2 - origin.a) */
  STG_RT_ANGLES_EXPLICIT = 2 /* headings[] supplied by the caller (single rays: fiducial line of
                                sight model_fiducial.cc:179, bumper, gripper) */
};

/* Model::Visibility bits the predicates read (stage.hh:1924-1937). */
#define STG_RT_VIS_GRIPPER_RETURN 1u
#define STG_RT_VIS_OBSTACLE_RETURN 2u
#define STG_RT_VIS_BLOB_RETURN 4u

/* stg_rt_config.flags */
#define STG_RT_CFG_DEFAULT 0u
/* Fiducial records: range and bearing (and the id cut-off at max_range_id) are computed once more on the host with
   the host C library's hypot / atan2 when the results are delivered, so that they agree with a CPU run of
   ModelFiducial::AddModelIfVisible (model_fiducial.cc:130,144,218) to the last bit; without it they are CUDA's
   (within 1e-9). Costs a pass over the delivered records on the calling thread: meant for libstage-sized worlds. */
#define STG_RT_CFG_HOST_EXACT_FIDUCIALS 1u

typedef struct stg_rt_ctx stg_rt_ctx;

typedef struct {
  uint32_t struct_size;   /* sizeof(stg_rt_config), for ABI growth */
  int32_t device;         /* CUDA device ordinal */
  double ppm;             /* World::ppm = 1/resolution (world.cc:406) */
  /* Initial extent of the dense device grid in SuperRegion units (1024 x 1024 cells each, region.hh:13-18):
     superregions [sreg_x0, sreg_x0+sreg_nx) x [sreg_y0, sreg_y0+sreg_ny). Rays may leave it (outside is empty
     space, as for a missing SuperRegion, world.cc:820-823). A polygon rendered outside it makes the grid GROW, as
     World::AddSuperRegion grows the reference's world (world.cc:1072-1093): the arrays are re-laid over the larger
     rectangle on the device (stg_rt_stats.extent_grows counts it; O(grid), so give the expected extent here).
     STG_RT_EEXTENT is left for worlds beyond 4095 SuperRegions. */
  int32_t sreg_x0, sreg_y0, sreg_nx, sreg_ny;
  uint32_t max_models;    /* capacity of the model table */
  uint32_t max_blocks;    /* capacity of the block table */
  uint32_t max_cell_entries; /* expected peak number of (cell, block) entries per layer; the
                                per-layer cell-entry table is sized to >= 2x this. 0 = 1<<22 */
  uint32_t flags;
  uint32_t max_delta_bytes; /* size of the moved-block delta buffer = the per-rank slot of the
                               all-gather; bounds the deltas of one flush round. 0 = 8 MiB */
  uint32_t reserved;
} stg_rt_config;

/* One sensor fan = one call of ModelRanger::Sensor::Update, one World::Raytrace(fan), or a group of
   explicit single rays sharing finder / origin / range. 64 bytes. */
typedef struct {
  uint32_t finder;        /* Ray::mod (stage.hh:718) */
  uint32_t n_samples;     /* sample_count / results.size() / number of explicit headings */
  uint8_t pred;           /* STG_RT_PRED_* */
  uint8_t ztest;          /* Ray::ztest */
  uint8_t layer;          /* layer to read: (World::updates + 1) % 2 at call time (world.cc:804) */
  uint8_t angles;         /* STG_RT_ANGLES_* */
  uint32_t first_heading; /* EXPLICIT: index of this fan's first heading in headings[] */
  double ox, oy, oz, oa;  /* global origin pose (Ray::origin; see STG_RT_ANGLES_*) */
  double range;           /* Ray::range (Sensor::range.max, fiducial max_range_anon, ...) */
  double fov;             /* RANGER / EVEN_FOV */
} stg_rt_fan;

/* Where the results of one submit go. All arrays are indexed by beam in fan order (fan 0's
   beams, then fan 1's, ...), length = sum of n_samples. NULL = not wanted (not computed or not
   copied back). Buffers stay owned by the caller and must remain valid until stg_rt_sync returns.
   Buffers obtained from stg_rt_host_alloc are written by the GPU directly (no staging copy). */
typedef struct {
  double *ranges;       /* RaytraceResult::range -> Sensor::ranges (model_ranger.cc:244-248) */
  double *intensities;  /* hit ? hit.vis.ranger_return : 0 -> Sensor::intensities (:250) */
  double *bearings;     /* RANGER: start_angle + t*sample_incr -> Sensor::bearings (:251);
                           others: the beam heading relative to origin.a */
  int32_t *hit_model;   /* RaytraceResult::mod as a model id, -1 = no hit */
  int32_t *hit_cell;    /* 2 x int32 per beam: global cell index of the hit cell */
  uint32_t *steps;      /* 2 x uint32 per beam: cell visits C, region probes R (diagnostics) */
} stg_rt_fan_out;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces: nothing in the reference (the CPU grid is created lazily, world.cc:1058-1093).
   Called from World::LoadWorldPostHook (world.cc:402) once ppm is known. */
int stg_rt_create(const stg_rt_config *cfg, stg_rt_ctx **out);
/* Called from atexit, NOT from ~World (which crashes in the reference). */
int stg_rt_destroy(stg_rt_ctx *ctx);
/* Text of the last error on this context (or of a failed create when ctx == NULL). */
const char *stg_rt_last_error(const stg_rt_ctx *ctx);
int stg_rt_abi_version(void);

/* ---- world state ------------------------------------------------------------------------ */

/* Publish the attributes of a model that the ray predicates read: Model::vis (stage.hh:1924-1937),
   the root of its tree (Model::IsRelated == same root, model.cc:446-466) and its colour
   (RaytraceResult::color, world.cc:853). Call at creation and whenever one changes
   (SetRangerReturn, SetFiducialReturn, SetParent, SetColor ...). rgba may be NULL. */
int stg_rt_model_upsert(stg_rt_ctx *ctx, uint32_t id, uint32_t root_id, double ranger_return,
                        int32_t fiducial_return, int32_t fiducial_key, uint32_t vis_flags,
                        const double *rgba);

/* Block::Map(layer), block.cc:170-181: xy = the n_pts pixel-space vertices Model::LocalToPixels
   produced (model.cc:474-491), i.e. the argument of World::MapPoly (world.cc:990); zmin/zmax =
   Block::global_z after the call (shared by both layers). Mapping a block that is already mapped
   on that layer renders it once more on top, as Block::Map does (it does not check; the reference's
   loader does this to bumper models): duplicate entries in the cells, counted by Region::count,
   all removed by the next UnMap (region.cc:292-299); the block keeps its place in the cell lists. */
int stg_rt_block_map(stg_rt_ctx *ctx, uint32_t block_id, uint32_t model_id, int layer, double zmin,
                     double zmax, int n_pts, const int32_t *xy);
/* Block::UnMap(layer), block.cc:183-189. Unmapping an unmapped block is a no-op, as there. */
int stg_rt_block_unmap(stg_rt_ctx *ctx, uint32_t block_id, int layer);

/* n times { stg_rt_block_unmap(blocks[i], layer); stg_rt_block_map(blocks[i], models[i], layer,
   z[2i], z[2i+1], first_pt[i+1]-first_pt[i], xy + 2*first_pt[i]); } in one call: what
   ModelPosition::Move does for every moving model each tick (model_position.cc:549-552), for
   callers that move many bodies at once. first_pt has n+1 entries. */
int stg_rt_blocks_remap(stg_rt_ctx *ctx, uint32_t n, const uint32_t *blocks, const uint32_t *models, int layer,
                        const double *z, const uint32_t *first_pt, const int32_t *xy);

/* ---- poses in, outlines on the device -------------------------------------------------------- */

/* Register the geometry of a block once, so that the device can render it from its model's pose: pts_xy = the
   n_pts polygon points in model coordinates exactly as Block::pts holds them after BlockGroup::CalcSize
   (blockgroup.cc:84-112; x0 y0 x1 y1 ...), local_zmin / local_zmax = Block::local_z. Blocks of one model are
   rendered in the order they were defined in (BlockGroup::Map's order, blockgroup.cc:114-121). A block defined again
   takes the new geometry at its next rendering; it must not be rendered from the old one at that moment. */
int stg_rt_block_define(stg_rt_ctx *ctx, uint32_t block_id, uint32_t model_id, double local_zmin, double local_zmax,
                        int n_pts, const double *pts_xy);

/* Model::UnMap(layer) + Model::Map(layer) (model.cc:493-519: every block of the model's BlockGroup) for n models at
   once, with the outlines derived ON THE DEVICE: gpose_xyza[4i..4i+3] = GetGlobalPose() + geom.pose of models[i],
   the frame Model::LocalToPixels (model.cc:474-491) composes every vertex on - pixel = floor((gpose + Pose(pt)) * ppm),
   with the host C library's cos / sin of gpose.a reproduced bit for bit - and z of which Block::Map adds to local_z
   (block.cc:177-180). What ModelPosition::Move does to the grid for every moving model each tick
   (model_position.cc:549-552) then costs 32 bytes per model from the caller and 48 per block in the moved-block deltas
   (stg_rt_delta_export), instead of a rasterised outline per block. Children are models of their own: list them too.
   A block rendered this way can still be unmapped, re-mapped with stg_rt_block_map, collision-tested ...: the library
   fetches its outline from the device when a call needs it on the host. */
int stg_rt_models_pose_update(stg_rt_ctx *ctx, uint32_t n, const uint32_t *models, const double *gpose_xyza, int layer);

/* ---- rays --------------------------------------------------------------------------------- */

/* Queue fans. Replaces the beam loops of ModelRanger::Sensor::Update (model_ranger.cc:232-255),
   World::Raytrace(fan) (world.cc:721-744, used by ModelBlobfinder::Update model_blobfinder.cc:170)
   and single World::Raytrace(Ray) calls (world.cc:756; ModelFiducial::AddModelIfVisible
   model_fiducial.cc:179). fans[] and headings[] are copied before the call returns.
   headings: one double per beam of the EXPLICIT fans (global heading, Ray::origin.a), else NULL.
   trig: NULL, or 2 doubles (cos, sin of the heading the reference would pass to cos()/sin(),
   world.cc:773-775) per beam of this submit in beam order - "strict" mode: use the host libm's
   values instead of the device's, for bit-for-bit agreement with a CPU run on that host. */
int stg_rt_fans_submit(stg_rt_ctx *ctx, uint32_t n_fans, const stg_rt_fan *fans,
                       const double *headings, const double *trig, const stg_rt_fan_out *out);

/* Sensor noise of one ranger fan (ModelRanger::Sensor::range_noise_const / range_noise / angle_noise,
   worldfile "noise [c p a]", model_ranger.cc:150, stage.hh:2637-2641). 32 bytes. */
typedef struct {
  double range_noise_const; /* VARIANCE of the additive Gaussian term, as generateGaussianNoise takes it */
  double range_noise;       /* proportional term: range * range_noise * U, U in {-1, -0.998 .. 0.998} */
  double angle_noise;       /* heading jitter: sample_incr * angle_noise * U * 0.5 */
  uint64_t seed;            /* stream selector; 0 = derive from the fan's finder id */
} stg_rt_fan_noise;

/* stg_rt_fans_submit for RANGER fans with sensor noise (model_ranger.cc:232-249): beam t is traced
   along (float)(a_t + sample_incr*angle_noise*U*0.5) and, if it hit closer than range.max, reports
   range + range*range_noise*U' + N(0, range_noise_const). The reference draws U, U', N from libc
   rand() in beam order, which cannot be replayed in parallel; here they come from a counter-based
   generator keyed by (seed, beam, launch), so the noise has the reference's DISTRIBUTION (same discrete
   uniform, same Box-Muller normal), not its values: statistical parity only (tests/test_gpu_noise.py).
   noise[i] belongs to fans[i]; all-zero entries give exactly the noise-free result. */
int stg_rt_fans_submit_noisy(stg_rt_ctx *ctx, uint32_t n_fans, const stg_rt_fan *fans,
                             const stg_rt_fan_noise *noise, const stg_rt_fan_out *out);

/* ---- fiducial finders ------------------------------------------------------------------------ */

/* A model with fiducial_return != 0 (World::models_with_fiducials, stage.hh:798): what
   ModelFiducial::AddModelIfVisible reads of "him" (model_fiducial.cc:109-224). 72 bytes. */
typedef struct {
  uint32_t model;           /* model id */
  int32_t fiducial_return;  /* him->vis.fiducial_return */
  int32_t fiducial_key;     /* him->vis.fiducial_key */
  uint32_t reserved;
  double x, y, z, a;        /* him->GetGlobalPose() */
  double sx, sy, sz;        /* him->GetGeom().size */
} stg_rt_fid_target;

/* One ModelFiducial::Update (model_fiducial.cc:229-293). 104 bytes. */
typedef struct {
  uint32_t finder;          /* model id of the fiducial sensor */
  int32_t fiducial_key;     /* this->vis.fiducial_key (:117) */
  uint8_t ignore_zloc;      /* ModelFiducial::ignore_zloc (:184) */
  uint8_t layer;            /* layer the line-of-sight rays read: (World::updates + 1) % 2 */
  uint8_t pad[6];
  double x, y, z, a;        /* GetGlobalPose() (mypose, :124) */
  double ox, oy, oz, oa;    /* GetGlobalPose() + geom.pose: the frame Model::LocalToGlobal composes
                               the ray pose on (stage.hh:2296; used by Model::Raytrace, :2003) */
  double max_range_anon, max_range_id, fov;
} stg_rt_fid_finder;

/* ModelFiducial::Fiducial (stage.hh:2555-2566). 88 bytes. */
typedef struct {
  uint32_t target;          /* index into the submitted targets[] (-> Fiducial::mod) */
  int32_t id;               /* Fiducial::id: fiducial_return, or 0 beyond max_range_id */
  double range, bearing;
  double geom[4];           /* size x, y, z and heading relative to the finder */
  double pose[4];           /* target's global pose */
} stg_rt_fiducial;

/* Replaces the body of ModelFiducial::Update for n_finders sensors at once: the candidate query
   (x/y window over the fiducial-bearing models), AddModelIfVisible's culls (key, range, field of
   view, relatedness), one line-of-sight ray per surviving candidate and the accept step. targets[]
   must be in the order the reference visits candidates (ascending Model*, :266-283); each finder's
   fiducials come back in that order: out[f * max_per_finder + k], k < min(out_count[f],
   max_per_finder). out / out_count are delivered by stg_rt_sync like fan results. */
int stg_rt_fiducials_submit(stg_rt_ctx *ctx, uint32_t n_targets, const stg_rt_fid_target *targets,
                            uint32_t n_finders, const stg_rt_fid_finder *finders,
                            uint32_t max_per_finder, stg_rt_fiducial *out, uint32_t *out_count);
/* The same with the records handed over packed: out[] holds finder 0's min(out_count[0], max_per_finder)
   records, then finder 1's, ... (a caller that walks the finders in order needs no stride, and at
   100,000 finders the strided array is 845 MB of mostly nothing). out_capacity = records out[] can
   hold; what does not fit is dropped from the end (the counts still tell). */
int stg_rt_fiducials_submit_packed(stg_rt_ctx *ctx, uint32_t n_targets, const stg_rt_fid_target *targets,
                                   uint32_t n_finders, const stg_rt_fid_finder *finders, uint32_t max_per_finder,
                                   uint64_t out_capacity, stg_rt_fiducial *out, uint32_t *out_count);

/* stg_rt_fiducials_submit_packed with the target records already in DEVICE memory - the output of the ranks' all-gather
   of their own robots' stg_rt_fid_target records (World::models_with_fiducials is world-wide, world.cc:623-629), which
   then never visits the host. Records whose model id is not below max_models are skipped: ranks pad their slots of the
   gathered array with them (model = 0xffffffff). The array must be complete in the order of the context's stream
   (stg_rt_stream_wait). Not available with STG_RT_CFG_HOST_EXACT_FIDUCIALS. */
int stg_rt_fiducials_submit_device_targets(stg_rt_ctx *ctx, uint32_t n_targets, const void *dev_targets, uint32_t n_finders,
                                           const stg_rt_fid_finder *finders, uint32_t max_per_finder, uint64_t out_capacity,
                                           stg_rt_fiducial *out, uint32_t *out_count);

/* ---- whole ranger models ---------------------------------------------------------------------- */

/* One ModelRanger::Sensor as the world file describes it (Sensor::Load, model_ranger.cc:129-156): pose
   relative to the ranger model, size (only z is read, :223), range.max, fov, sample count. 64 bytes. */
typedef struct {
  double px, py, pz, pa;
  double size_z, range_max, fov;
  uint32_t n_samples, reserved;
} stg_rt_ranger_sensor;

/* One ranger model this tick: x, y, z, a = GetGlobalPose() + geom.pose, the frame Model::LocalToGlobal
   composes a sensor's pose on (stage.hh:2296). 40 bytes. */
typedef struct {
  uint32_t finder;          /* the ranger model's id */
  uint8_t layer;            /* (World::updates + 1) % 2 */
  uint8_t pad[3];
  double x, y, z, a;
} stg_rt_ranger_pose;

/* ModelRanger::Update (model_ranger.cc:202-209) for n_rangers ranger models that share one sensor list
   (16 sonars on every pioneer): what stg_rt_fans_submit does for n_rangers * n_sensors fans, with the
   fans derived on the device - origin of the first ray = LocalToGlobal(sensor pose + start angle,
   + size.z / 2) (:218-226; cos / sin of the model's heading as the host C library gives them) - so
   that 40 bytes per model cross PCIe instead of 64 per sensor. Beams come back model by model,
   sensor by sensor: beam t of sensor s of model r is at (r * B + first(s) + t), B = samples per
   model. Must be the only submit of its batch (it flushes what is queued). */
int stg_rt_rangers_submit(stg_rt_ctx *ctx, uint32_t n_sensors, const stg_rt_ranger_sensor *sensors, uint32_t n_rangers,
                          const stg_rt_ranger_pose *poses, const stg_rt_fan_out *out);

/* ---- collision tests ------------------------------------------------------------------------ */

/* Model::TestCollision (model.cc:666-679) for n_queries models at once, against the device grid as
   it stands after every Map / UnMap issued so far. Query q covers blocks[query_first[q] ..
   query_first[q + 1]): the model's own blocks in BlockGroup order, then its descendants', depth
   first - the order Model::TestCollision and BlockGroup::TestCollision (blockgroup.cc:41-50) visit
   them in. layer = World::updates % 2, the layer Block::TestCollision reads (block.cc:140).
   hit_model[q] = id of the model the reference would return: the first block found, walking the
   query's blocks, their rendered cells and those cells' block lists in order, that belongs to
   another, unrelated model with obstacle_return and overlaps in z (block.cc:143-160);
   ground_model (World::GetGround()'s id) for a block that reaches below z = 0 (:137-138); -1 for no
   collision. Blocks of models without obstacle_return test nothing (:136). Delivered by
   stg_rt_sync, like sensor results. */
int stg_rt_collisions_test(stg_rt_ctx *ctx, int layer, uint32_t n_queries, const uint32_t *query_first,
                           const uint32_t *blocks, uint32_t ground_model, int32_t *hit_model);

/* Model::AppendTouchingModels (model.cc:661-664 -> BlockGroup::AppendTouchingModels blockgroup.cc:35-39 ->
   Block::AppendTouchingModels block.cc:116-127; called by Model::UpdateCharge, model.cc:700-701, for every
   model that gives charge) for n_queries models at once. Query q covers the model's OWN blocks
   blocks[query_first[q] .. query_first[q + 1]) (children are not visited). layer = World::updates % 2
   (block.cc:118). touchers[q * max_per_query ..] = ids of the models, not related to the query's (other
   tree), that have a block in any cell those blocks are rendered into - no obstacle_return, no z test -
   ascending; counts[q] = how many there are (the list is cut at max_per_query; 0xffffffff: more than
   the 120 distinct models one query can hold). Delivered by stg_rt_sync, like sensor results. */
int stg_rt_touching_models(stg_rt_ctx *ctx, int layer, uint32_t n_queries, const uint32_t *query_first,
                           const uint32_t *blocks, uint32_t max_per_query, uint32_t *touchers, uint32_t *counts);

/* ---- batched ModelPosition::Move ------------------------------------------------------------ */

/* ModelPosition::Move (model_position.cc:526-566) for n_movers position models at once, with the
   result the reference's one-by-one loop (world.cc:647-648) produces: mover i is re-rendered at the
   pose it wants (UnMapWithChildren / MapWithChildren on `layer` = World::updates % 2), tested with
   Model::TestCollision against everything as it stands THEN - the movers before it where they ended
   up, the movers after it where the layer still has them - and, if it hit something, re-rendered at
   the pose it has now (stalled[i] = 1, Model::SetStall). Mover i covers blocks[mover_first[i] ..
   mover_first[i + 1]): its own blocks and its descendants', depth first. Per block: z / first_pt /
   xy = z range and outline at the wanted pose, z_start / first_pt_start / xy_start = at the present
   pose (startpose, :541), both as Model::LocalToPixels gives them. The movers must come in the order
   the reference moves them (World::active_velocity) and belong to different model trees.
   How: what the wanted outlines touch is collected against the layer as it is and against the layer
   with every mover at its wanted pose, and what the present outlines touch against the latter; a
   mover that touches nothing, or something that does not move, is decided right there, and the few
   that touch other movers are decided in order from those three lists. The cell lists end up in the
   reference's order (each block's last Map keeps its place in the loop's sequence). Synchronous;
   single process (not combined with stg_rt_delta_export). */
int stg_rt_models_move(stg_rt_ctx *ctx, int layer, uint32_t n_movers, const uint32_t *mover_first, const uint32_t *blocks,
                       const double *z, const uint32_t *first_pt, const int32_t *xy, const double *z_start,
                       const uint32_t *first_pt_start, const int32_t *xy_start, uint8_t *stalled);

/* ---- blobfinders ----------------------------------------------------------------------------- */

/* One ModelBlobfinder::Update (model_blobfinder.cc:165-250). 72 bytes. */
typedef struct {
  uint32_t finder;          /* model id of the blobfinder */
  uint32_t scan_width, scan_height;
  uint8_t layer, pad[3];
  uint32_t first_color, n_colors; /* this sensor's tracked colours in colors_rgb[]; 0 = all */
  double ox, oy, oz, oa;    /* LocalToGlobal(Pose(0,0,0,pan)): global pose of the scan centre */
  double range, fov;
} stg_rt_blob_finder;

/* ModelBlobfinder::Blob (stage.hh:2333-2338). 32 bytes. Blob::color = colour of `model`. */
typedef struct {
  uint32_t model;           /* the model seen by the blob's first sample */
  uint32_t left, top, right, bottom, pad;
  double range;
} stg_rt_blob;

/* Replaces ModelBlobfinder::Update for n_finders sensors: the scan_width-ray fan (no z test, stop
   at any unrelated model) and the run-length colour segmentation. Model colours are the rgba given
   to stg_rt_model_upsert. colors_rgb = 3 doubles per tracked colour. Results as for fiducials. */
int stg_rt_blobs_submit(stg_rt_ctx *ctx, uint32_t n_finders, const stg_rt_blob_finder *finders,
                        uint32_t n_colors, const double *colors_rgb, uint32_t max_per_finder,
                        stg_rt_blob *out, uint32_t *out_count);

/* Start everything queued (deltas, then fans) on the GPU without waiting. Implicitly done by
   stg_rt_sync, and by a map/unmap issued while fans are still queued (ordering rule above). */
int stg_rt_flush(stg_rt_ctx *ctx);
/* Wait for all started work and deliver results into the stg_rt_fan_out buffers. Called from
   World::Update after the worker threads finished (world.cc:650-657), before :668. */
int stg_rt_sync(stg_rt_ctx *ctx);

/* Page-locked host memory the GPU can write results into directly. */
void *stg_rt_host_alloc(stg_rt_ctx *ctx, size_t bytes);
int stg_rt_host_free(stg_rt_ctx *ctx, void *p);

/* ---- resident fan sets: inputs and outputs stay in HBM ------------------------------------ */
/* For callers that keep sensor geometry on the device across ticks (and for measuring the kernels
   alone). A set is a batch of fans uploaded once; stg_rt_set_trace re-traces it against the current
   grid; results stay on the device until stg_rt_set_download. */
typedef struct stg_rt_set stg_rt_set;
int stg_rt_set_create(stg_rt_ctx *ctx, uint32_t n_fans, const stg_rt_fan *fans, const double *headings,
                      const double *trig, uint32_t want /* STG_RT_WANT_* */, stg_rt_set **out);
int stg_rt_set_destroy(stg_rt_ctx *ctx, stg_rt_set *set);
/* Replace the origin pose / layer of every fan of the set (robots moved): n_fans entries. */
int stg_rt_set_update_origins(stg_rt_ctx *ctx, stg_rt_set *set, const double *xyza, int layer);
/* Asynchronous: applies queued deltas first, then traces. */
int stg_rt_set_trace(stg_rt_ctx *ctx, stg_rt_set *set);
int stg_rt_set_download(stg_rt_ctx *ctx, stg_rt_set *set, const stg_rt_fan_out *out);
/* Device pointers of the set's result arrays (NULL members for arrays not in `want`). */
int stg_rt_set_device_ptrs(stg_rt_ctx *ctx, stg_rt_set *set, stg_rt_fan_out *dev);
uint64_t stg_rt_set_beams(const stg_rt_set *set);

#define STG_RT_WANT_RANGES 1u
#define STG_RT_WANT_INTENSITIES 2u
#define STG_RT_WANT_BEARINGS 4u
#define STG_RT_WANT_HIT_MODEL 8u
#define STG_RT_WANT_HIT_CELL 16u
#define STG_RT_WANT_STEPS 32u

/* ---- grid inspection (parity tests; World::Raytrace debug aids rt_cells, stage.hh:873) ------ */

/* Copy out every (cell, block) entry of the device grid as int32[5] records
   {layer, x, y, seq_rank, block}: seq_rank = position of the entry in the reference's per-cell
   list order (Cell::blocks[layer], region.hh:47). Returns the entry count (may exceed cap), < 0 on
   error. Also Region::count (both layers, region.cc:19-34) as int64[3] {rx, ry, count} records. */
int64_t stg_rt_grid_dump(stg_rt_ctx *ctx, int32_t *records, int64_t cap, int64_t *region_counts,
                         int64_t counts_cap, int64_t *n_counts);

/* A digest of the whole device grid without moving it: out[0], out[1] = per layer, the sum over every distinct
   (cell, block) entry of a 64-bit mix of (x, y, block, position of the block in Cell::blocks[layer] order); out[2],
   out[3] = the number of such entries per layer; out[4] = the same kind of sum over (region, Region::count). Two grids
   agree on all five iff they hold the same blocks in the same order in every cell (up to 2^-64). Replicas of a
   multi-GPU run compare it with each other; the tests compare it with the digest a CPU restatement of the reference takes of its own lists.
   Synchronous. */
int stg_rt_grid_hash(stg_rt_ctx *ctx, uint64_t out[5]);

/* The pixel outline block `block_id` is rendered with on `layer` right now (as World::MapPoly received it), whether the
   caller supplied it or the device derived it from a pose: up to cap_pts points into xy (2 x int32 each). Returns the
   number of points (0: not mapped). Synchronous; for parity tests of the pose path against Model::LocalToPixels. */
int stg_rt_debug_block_outline(stg_rt_ctx *ctx, uint32_t block_id, int layer, int32_t *xy, int cap_pts);

/* The region owner tags the march uses to walk through regions that hold nothing but the finder's own
   tree (no counterpart in the reference: an acceleration that must never change a result; tests check
   its invariant against the grid). int64[3] {rx, ry, tag} per region with a non-zero tag; tag = root
   model id + 1, or 0xffffffff for "more than one tree". Returns the record count (may exceed cap). */
int64_t stg_rt_debug_region_owners(stg_rt_ctx *ctx, int64_t *records, int64_t cap);

/* The occupancy tiles and their summary words (what the march reads instead of Cell::blocks, region.hh:47) against the
   cell entries they stand for: out[0] = cells whose bit in a row or column mask differs from "the cell holds an entry",
   out[1] = summary bits that differ from "that row / column mask is non-zero", out[2] = occupied cells seen (both
   layers). No counterpart in the reference: accelerations that must never change a result; the tests expect zeros
   after every kind of update. Synchronous. */
int stg_rt_debug_occupancy_check(stg_rt_ctx *ctx, uint64_t out[3]);

/* ---- multi-GPU: moved-block deltas ---------------------------------------------------------- */
/* Robots are sharded over ranks (one process per GPU), the grid is replicated. Each tick every rank
   serialises the deltas its own robots produced, the ranks all-gather the byte blobs (NCCL over
   NVLink, done by the caller: torch.distributed / ncclAllGather), and every rank applies the other
   ranks' blobs. Blobs are applied in rank order so that all replicas, and a single-GPU run that
   issues the same maps in rank order, hold identical grids. */

/* Move the deltas queued since the last flush into a device buffer owned by the context and return
   its address and size (fixed-size records). They are NOT applied locally yet.
   ORDERING: the blob is written by an asynchronous copy on the context's stream (stg_rt_get_stream): it is
   complete for work queued LATER ON THAT STREAM. A caller whose collective runs on another stream (its own
   ncclAllGather stream) must call stg_rt_stream_signal(ctx, that_stream) after the export and
   stg_rt_stream_wait(ctx, that_stream) after enqueuing the collective, before stg_rt_delta_apply_gathered -
   or hand the context its stream with stg_rt_set_stream. K1 refuses a slot that does not hold a well-formed
   blob (magic word, record counts that add up to the blob's own size and fit the slot) and the next
   stg_rt_sync then fails with STG_RT_EINVAL, so a gather that ran too early cannot corrupt the grid silently. */
int stg_rt_delta_export(stg_rt_ctx *ctx, int rank, void **dev_blob, size_t *n_bytes);
/* Everything queued so far on the context's stream happens before work queued later on consumer_stream. */
int stg_rt_stream_signal(stg_rt_ctx *ctx, void *consumer_stream);
/* Everything queued so far on producer_stream happens before work queued later on the context's stream. */
int stg_rt_stream_wait(stg_rt_ctx *ctx, void *producer_stream);
/* Size every rank must use as the per-rank slot of the all-gather buffer. */
size_t stg_rt_delta_slot_bytes(const stg_rt_ctx *ctx);
/* Apply n_ranks blobs laid out back to back, slot_bytes apart, in a DEVICE buffer (the all-gather
   output), in rank order, directly out of that buffer. Includes this rank's own blob. Runs on the
   context's stream: the gathered buffer must be complete in that stream's order (see stg_rt_delta_export).
   Fails with STG_RT_ENOMEM when the cell-entry tables would pass half full (as a local flush does). */
int stg_rt_delta_apply_gathered(stg_rt_ctx *ctx, const void *dev_gathered, int n_ranks, size_t slot_bytes);

/* ---- plumbing -------------------------------------------------------------------------------- */
/* Run all work of this context on the given cudaStream_t (e.g. torch's current stream) instead of
   the context's own stream. NULL restores the own stream. */
int stg_rt_set_stream(stg_rt_ctx *ctx, void *cuda_stream);
void *stg_rt_get_stream(stg_rt_ctx *ctx);

/* Diagnostics: (cos, sin) of n headings exactly as the ray march computes them on the device
   (World::Raytrace's zero-heading substitution included, world.cc:773-775). cos_sin = 2n doubles.
   Synchronous. Used to check the device trig against the host C library. */
int stg_rt_debug_trig(stg_rt_ctx *ctx, uint64_t n, const double *headings, double *cos_sin);
/* Diagnostics: the per-beam headings the device derives for these fans (the float-rounded running
   angle of ModelRanger::Sensor::Update, model_ranger.cc:237-241,254; the even spacing of
   World::Raytrace(fan), world.cc:731-743), sum(n_samples) doubles in fan order. Synchronous. */
int stg_rt_debug_headings(stg_rt_ctx *ctx, uint32_t n_fans, const stg_rt_fan *fans, const double *headings_in,
                          double *headings_out);

/* ---- libc rand() stream of a ranger whose beam loop no longer runs on the host ------------------- */
/* ModelRanger::Sensor::Update draws from libc rand() inside its beam loop even when every noise parameter is
   zero (model_ranger.cc:232-249: one simpleNoise() per beam for the heading jitter, :237; for a beam that
   returned less than range.max one more simpleNoise() and one generateGaussianNoise(), :244-246, which itself
   draws two numbers on every second call, :181-199 - all multiplied by 0). Controllers that share the stream
   (examples/ctrl/lasernoise.cc) see different numbers once the loop is gone. This advances rand() by exactly the
   draws the loop would have made for one sensor: n_samples beams of which n_in_range came back shorter than
   range.max. It mirrors generateGaussianNoise's static "spare" flag with one of its own, so every ranger sensor of
   the process has to go through it (or none). Pure host code; callable without a context. Returns the number of
   rand() calls made. */
uint32_t stg_rt_ranger_rand_advance(uint32_t n_samples, uint32_t n_in_range);

/* Diagnostics: what the link from this GPU to page-locked host memory sustains, in GB/s, over `reps` transfers of
   `bytes`: mode 0 = stores from a kernel straight into host memory (how the march delivers ranges and intensities),
   mode 1 = cudaMemcpyAsync device-to-host. Run on several GPUs at once (tools/link_ceiling.py) it shows the ceiling the
   host's PCIe / memory system puts on N result streams. Synchronous. */
int stg_rt_debug_link_bandwidth(stg_rt_ctx *ctx, int mode, size_t bytes, int reps, double *gb_per_s);

/* Counters since create: kernel launches by kind, rays traced, bytes copied each way. */
typedef struct {
  uint64_t launches_rasterise, launches_march, launches_post, launches_other;
  uint64_t beams, fans, map_calls, unmap_calls, edge_steps;
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t extent_grows; /* times the device grid was re-laid over a larger extent (World::AddSuperRegion) */
} stg_rt_stats;
int stg_rt_get_stats(stg_rt_ctx *ctx, stg_rt_stats *out);

/* Device time in milliseconds of the most recent launch of each kernel group, measured with CUDA
   events recorded on the context's stream around the launches: ms[0] = delta application (K1),
   ms[1] = fan headings (K2a), ms[2] = ray march with fused epilogue (K2+K3). -1 = never launched.
   Waits for the stream. */
int stg_rt_kernel_times(stg_rt_ctx *ctx, float ms[3]);
/* The same for the first n of: [0] delta application (K1), [1] fan headings (K2a), [2] ray march (K2 + ranger
   epilogue), [3] fiducial finders (K3: candidate query, line-of-sight rays, packing), [4] blobfinders (scan fan +
   segmentation), [5] collision tests / touching models (K4, K4b). -1 = never launched. */
int stg_rt_kernel_times_ex(stg_rt_ctx *ctx, float *ms, int n);

#ifdef __cplusplus
}
#endif
#endif /* STG_RT_H */
