/* Event-trace format shared by the reference harness (oracle/ref_harness.cc, which WRITES traces
 * by interposing the unmodified reference), the T1 restatement (oracle/t1_oracle.cc, which
 * REPLAYS them) and the python reader (tests/trace_io.py). TEST INFRASTRUCTURE ONLY.
 *
 * A trace is everything the raycast hot path saw during a run of the reference, in call order:
 *   MAP    Block::Map(layer)      block.cc:170-181  -> pixel-space polygon + global z bounds
 *   UNMAP  Block::UnMap(layer)    block.cc:183-189
 *   RAYS   World::Raytrace(Ray)   world.cc:756-963  -> inputs and the reference's answer
 *   TICK   World::Update finished world.cc:605-676
 *   RANGER_SCAN  ModelRanger::Sensor::Update model_ranger.cc:211-256 -> the fan's inputs and the
 *          ranges / intensities / bearings it left in the Sensor (logged AFTER its RAYS)
 * Layout: TraceHeader, TraceModel[n_models], TraceEvent[n_events], int32 verts[2*n_verts],
 * TraceRay[n_rays], double aux[n_aux]. Little-endian, natural alignment, no padding surprises (all structs are
 * multiples of 8 bytes).
 */
#ifndef STG_TRACE_FORMAT_H
#define STG_TRACE_FORMAT_H
#include <stdint.h>

#define STG_TRACE_MAGIC "STGTRC05"

enum { TRACE_EV_MAP = 1, TRACE_EV_UNMAP = 2, TRACE_EV_RAYS = 3, TRACE_EV_TICK = 4,
       TRACE_EV_RANGER_SCAN = 5, TRACE_EV_FIDUCIAL = 6, TRACE_EV_BLOB = 7, TRACE_EV_COLLISION = 8,
       TRACE_EV_TOUCHERS = 9 /* added without a new magic: older traces simply hold none */ };

/* predicate kinds: which of the reference's static ray_test_func_t a ray was traced with */
enum {
  TRACE_KIND_RANGER = 0,    /* ranger_match              model_ranger.cc:158-170 */
  TRACE_KIND_UNRELATED = 1, /* fiducial_raytrace_match / blob_match / bumper_match
                               model_fiducial.cc:103-107, model_blobfinder.cc:102-107,
                               model_bumper.cc:157-162 */
  TRACE_KIND_GRIPPER = 2    /* gripper_raytrace_match    model_gripper.cc:329-336 */
};

#define TRACE_FLAG_GRIPPER_RETURN 1u
#define TRACE_FLAG_OBSTACLE_RETURN 2u
#define TRACE_FLAG_BLOB_RETURN 4u

typedef struct {
  char magic[8];
  double ppm;
  uint32_t n_models, n_events, n_verts, n_rays;
  uint32_t n_ticks, n_aux;
} TraceHeader; /* 40 B */

typedef struct {
  uint32_t id;     /* Model::id (stage.hh:1794) */
  uint32_t parent; /* 0xffffffff for a top-level model */
  uint32_t root;   /* id of the top of this model's tree; IsRelated == same root (model.cc:446-466) */
  int32_t fiducial_return, fiducial_key;
  uint32_t flags;
  double ranger_return;
  double color[4];
  char type[16];
} TraceModel; /* 80 B */

typedef struct {
  uint8_t type, layer;
  uint16_t n_pts;  /* MAP: polygon vertex count */
  uint32_t block;  /* MAP/UNMAP: dense block id (order of first appearance); TICK: World::updates */
  uint32_t model;  /* MAP: owning model id; RAYS: ray count */
  uint32_t first;  /* MAP: first vertex index; RAYS: first ray index; RANGER_SCAN: first aux index */
  double zmin, zmax; /* MAP: Block::global_z after the call */
} TraceEvent; /* 32 B */

typedef struct {
  double ox, oy, oz, oa; /* Ray::origin (global) */
  double range;          /* Ray::range */
  double res_range;      /* RaytraceResult::range */
  int32_t hit_model;     /* RaytraceResult::mod->id, or -1 */
  uint32_t finder;       /* Ray::mod->id */
  uint8_t kind, ztest, layer, pad[5];
} TraceRay; /* 64 B */

/* RANGER_SCAN: event.model = ranger model id, event.block = sensor index, event.n_pts = sample
 * count n, event.layer = layer the rays read, aux[first..] = { ox, oy, oz, oa (the global pose of
 * the first ray, rayorg after LocalToGlobal, model_ranger.cc:226-230), range_max, fov,
 * ranges[n], intensities[n], bearings[n] }. */
#define TRACE_SCAN_HEADER_DOUBLES 6

/* FIDUCIAL: one ModelFiducial::Update (model_fiducial.cc:229-293). event.model = finder model id,
 * event.layer = layer its rays read, event.block = number of fiducials found, aux[first..] =
 *   { x, y, z, a (GetGlobalPose), ox, oy, oz, oa (GetGlobalPose + geom.pose), max_range_anon,
 *     max_range_id, fov, fiducial_key, ignore_zloc, n_targets }                      14 doubles
 *   n_targets x { model id, fiducial_return, fiducial_key, x, y, z, a, sx, sy, sz }  10 doubles each,
 *     World::models_with_fiducials in ascending Model* order (the order candidates are visited)
 *   n_found  x { model id, id, range, bearing, geom[4], pose[4] }                    12 doubles each */
#define TRACE_FID_HEADER_DOUBLES 14
#define TRACE_FID_TARGET_DOUBLES 10
#define TRACE_FID_RESULT_DOUBLES 12

/* BLOB: one ModelBlobfinder::Update (model_blobfinder.cc:165-250). event.model = finder model id,
 * event.layer, event.block = number of blobs, aux[first..] =
 *   { ox, oy, oz, oa (LocalToGlobal(Pose(0,0,0,pan))), range, fov, scan_width, scan_height,
 *     n_colors }                                                                       9 doubles
 *   n_colors x { r, g, b }
 *   n_blobs  x { r, g, b, left, top, right, bottom, range }                            8 doubles each */
#define TRACE_BLOB_HEADER_DOUBLES 9
#define TRACE_BLOB_RESULT_DOUBLES 8

/* COLLISION: one top-level Model::TestCollision() (model.cc:666-679; ModelPosition::Move calls it
 * right after re-mapping the model, model_position.cc:549-554). event.model = the model tested,
 * event.layer = World::updates % 2 (the layer Block::TestCollision reads, block.cc:140),
 * event.block = id of the model it returned + 1, 0 for no collision, event.n_pts = n, aux[first ..
 * first + n) = the dense ids of the blocks it may visit, in visiting order: the model's own
 * BlockGroup, then each child's, depth first (blockgroup.cc:41-50, model.cc:668-676). */

/* TOUCHERS: one Model::AppendTouchingModels (model.cc:661-664, called by Model::UpdateCharge :700-701 for
 * every model that gives charge, each tick) -> BlockGroup / Block::AppendTouchingModels (blockgroup.cc:35-39,
 * block.cc:116-127): the set of models not related to this one that have a block in any cell this model's own
 * blocks are rendered into. event.model = the model, event.layer = World::updates % 2 (block.cc:118),
 * event.n_pts = n own blocks, event.block = m models found, aux[first .. first + n) = the dense ids of the
 * model's BlockGroup (children are NOT visited), aux[first + n .. first + n + m) = the ids of the models
 * added to the set, ascending. */

#endif
