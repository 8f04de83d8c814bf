// libstage_gpu.so - binds an UNMODIFIED libstage (rtv/Stage 4.3.0) to libstg_rt.so (include/stg_rt.h).
//
// libstage is a shared library built -fPIC with default symbol visibility, so its internal calls to
// Block::Map, Block::UnMap, ModelRanger::Sensor::Update, ModelFiducial::Update, ModelBlobfinder::Update and
// World::CallUpdateCallbacks go through the PLT. This library defines those six symbols itself; loaded
// before libstage (LD_PRELOAD=libstage_gpu.so stage -g worlds/simple.world, or linked ahead of it) it is
// bound first, does the GPU side, and forwards to libstage's own definition where the CPU side is still
// wanted. No header, no source file and no binary of Stage changes: controllers, ctrl plugins and
// libstageplugin keep working bit for bit (tests/test_gpu_libstage.py runs simple.world and fasr.world with
// their real controllers both ways and compares every ranger double, fiducial record and robot pose).
// INTEGRATION.md lists the same hooks as the patch a maintainer would apply in-tree instead.
//
// What runs where when STAGE_GPU=1:
//   Block::Map / UnMap (block.cc:170-189)          libstage's own (the CPU grid stays authoritative for
//                                                  Block::TestCollision, block.cc:129-168) + stg_rt_block_map/unmap
//   ModelRanger::Sensor::Update (model_ranger.cc:211-256)   set-up as there, beam loop replaced by stg_rt_fans_submit
//   ModelFiducial::Update (model_fiducial.cc:229-293)       candidate search, culls and line-of-sight rays replaced by
//                                                  stg_rt_fiducials_submit (all finders of a tick in one batch)
//   ModelBlobfinder::Update (model_blobfinder.cc:165-250)   scan fan and colour segmentation replaced by stg_rt_blobs_submit
//                                                  (all blobfinders of a tick in one batch; a blob's colour is the colour of
//                                                  the model its first sample hit, world.cc:854)
//   World::CallUpdateCallbacks (world.cc:668)      first stg_rt_sync: ranges / intensities / fiducials / blobs are in the
//                                                  sensors' std::vectors before any controller callback reads them
// Everything else (bumper, gripper rays, collisions) keeps calling World::Raytrace on the CPU grid, which
// this library never lets go stale. Any error from the C ABI is printed once, the context is dropped, the tick is
// redone with libstage's own loops and the process carries on as plain Stage (include/stg_rt.h: no CPU path INSIDE
// libstg_rt.so; the fallback is the caller's, as SURVEY.md 8b specifies).
//
// Build: integration/build.sh (needs the Stage headers; nothing of Stage is copied into this repository).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <dlfcn.h>

// Block::pts / group / global_z, Model::vis / parent / world, World::models ... are private or protected
#define private public
#define protected public
#include "stage.hh"
#undef private
#undef protected

#include "stg_rt.h"

using namespace Stg;

namespace {

template <class F> F next_symbol(const char *mangled)
{
  void *p = dlsym(RTLD_NEXT, mangled);
  if (!p) {
    fprintf(stderr, "libstage_gpu: libstage does not define %s (load this library BEFORE libstage)\n", mangled);
    abort();
  }
  return (F)p;
}

struct ModelSig { // what stg_rt_model_upsert was last told about a model
  uint32_t root;
  double ranger_return;
  int fiducial_return, fiducial_key;
  uint32_t flags;
  double rgba[4];
  bool operator==(const ModelSig &o) const { return !memcmp(this, &o, sizeof(*this)); }
};

struct MapCall {
  bool unmap;
  uint32_t block, model;
  int layer;
  double zmin, zmax;
  std::vector<int32_t> xy;
};

struct RangerJob {
  ModelRanger::Sensor *sensor;
  ModelRanger *mod;
  // STAGE_GPU=2 (verify): the GPU's answer goes here and is compared with what libstage's own loop put into the sensor
  std::vector<double> *gpu_ranges, *gpu_intensities;
};

// libc's random() / rand() state, saved and put back around a call whose draws must not count (emulation mode)
struct RngSnapshot {
  char bytes[128];
};
char *rng_swap_area()
{
  static char other[128];
  static bool ready = false;
  if (!ready) {
    char *cur = initstate(1u, other, sizeof(other)); // makes `other` current and returns the state in use so far
    setstate(cur);
    ready = true;
  }
  return other;
}
void rng_save(RngSnapshot &s)
{
  char *other = rng_swap_area();
  char *cur = setstate(other); // leaving a state writes its position into its first word; what comes back is that array
  memcpy(s.bytes, cur, sizeof(s.bytes));
  setstate(cur);
}
void rng_restore(const RngSnapshot &s)
{
  char *other = rng_swap_area();
  char *cur = setstate(other);
  memcpy(cur, s.bytes, sizeof(s.bytes));
  setstate(cur);
}

struct FidBatch { // finders that saw the same target list (normally all finders of a tick)
  std::vector<stg_rt_fid_target> targets;
  std::vector<Model *> target_models;
  std::vector<stg_rt_fid_finder> finders;
  std::vector<ModelFiducial *> owners;
  std::vector<stg_rt_fiducial> out;
  std::vector<uint32_t> count;
  uint32_t max_per_finder = 0;
  std::vector<std::vector<ModelFiducial::Fiducial> > expected; // verify: what libstage's own Update found, per finder
};

struct BlobJob { // one ModelBlobfinder::Update of the tick
  ModelBlobfinder *owner;
  stg_rt_blob_finder rec;
  std::vector<double> colors;                      // this finder's tracked colours, rgb
  std::vector<ModelBlobfinder::Blob> expected;     // verify / emulate: what libstage's own Update found
};

struct Shim {
  std::mutex mu;
  bool wanted = false, failed = false, verbose = false;
  bool verify = false; // STAGE_GPU=2: run libstage's own loops too and compare every answer in place
  // STAGE_GPU=3: no GPU at all - every answer is taken from libstage's own loops on the side (their rand() draws undone),
  // held back until the sync point and delivered there like a GPU answer, rand() replay included. Exercises everything
  // of this binding but the C ABI calls, on any machine (tests/test_libstage_binding_cpu.py).
  bool emulate = false;
  uint64_t verified_beams = 0, verified_fiducials = 0, mismatched_beams = 0, mismatched_fiducial_updates = 0;
  double fiducial_max_err = 0;
  stg_rt_ctx *ctx = nullptr;
  std::vector<MapCall> early;                        // Map / UnMap calls seen before the context exists
  bool early_open = true;                            // ... are still being collected (emulation: same allocations as a GPU run)
  std::unordered_map<const Block *, uint32_t> block_ids;
  std::unordered_map<const Model *, ModelSig> published;
  uint64_t published_at_update = ~0ull;
  std::vector<RangerJob> rangers;
  std::vector<FidBatch> fid;
  std::vector<BlobJob> blobs;
  uint64_t blob_updates = 0, verified_blobs = 0, mismatched_blob_updates = 0;
  uint64_t fans = 0, fid_updates = 0, maps = 0, ticks = 0;

  Shim()
  {
    const char *e = getenv("STAGE_GPU");
    wanted = e && atoi(e) != 0;
    verify = e && atoi(e) == 2;
    emulate = e && atoi(e) == 3;
    verbose = getenv("STAGE_GPU_VERBOSE") != nullptr;
  }
  bool active() const { return wanted && !failed; }
} shim;

void give_up(const char *what, int rc)
{
  if (!shim.failed)
    fprintf(stderr, "libstage_gpu: %s failed (%d: %s); continuing on the CPU path\n", what, rc,
            stg_rt_last_error(shim.ctx) ? stg_rt_last_error(shim.ctx) : "");
  shim.failed = true;
  if (shim.ctx) {
    stg_rt_destroy(shim.ctx);
    shim.ctx = nullptr;
  }
}

ModelSig signature(const Model *m)
{
  ModelSig s;
  memset(&s, 0, sizeof(s));
  s.root = const_cast<Model *>(m)->Root()->id;
  s.ranger_return = m->vis.ranger_return;
  s.fiducial_return = m->vis.fiducial_return;
  s.fiducial_key = m->vis.fiducial_key;
  s.flags = (m->vis.gripper_return ? STG_RT_VIS_GRIPPER_RETURN : 0u) | (m->vis.obstacle_return ? STG_RT_VIS_OBSTACLE_RETURN : 0u) |
            (m->vis.blob_return ? STG_RT_VIS_BLOB_RETURN : 0u);
  s.rgba[0] = m->color.r;
  s.rgba[1] = m->color.g;
  s.rgba[2] = m->color.b;
  s.rgba[3] = m->color.a;
  return s;
}

// mu held
bool publish_locked(const Model *m)
{
  const ModelSig s = signature(m);
  std::unordered_map<const Model *, ModelSig>::iterator it = shim.published.find(m);
  if (it != shim.published.end() && it->second == s)
    return true;
  const int rc = stg_rt_model_upsert(shim.ctx, m->id, s.root, s.ranger_return, s.fiducial_return, s.fiducial_key, s.flags, s.rgba);
  if (rc) {
    give_up("stg_rt_model_upsert", rc);
    return false;
  }
  shim.published[m] = s;
  return true;
}

// mu held. Attribute changes made by controllers (SetFiducialReturn, SetColor, SetParent ...) since the last tick.
bool publish_changed_locked(World *w)
{
  if (shim.emulate || shim.published_at_update == w->updates)
    return true;
  shim.published_at_update = w->updates;
  for (std::set<Model *>::const_iterator it = w->models.begin(); it != w->models.end(); ++it)
    if (!publish_locked(*it))
      return false;
  return true;
}

// mu held
uint32_t block_id_locked(const Block *b)
{
  std::unordered_map<const Block *, uint32_t>::iterator it = shim.block_ids.find(b);
  if (it != shim.block_ids.end())
    return it->second;
  const uint32_t id = (uint32_t)shim.block_ids.size();
  shim.block_ids[b] = id;
  return id;
}

// mu held
bool forward_map_locked(const MapCall &c, const Model *m)
{
  int rc;
  if (c.unmap) {
    rc = stg_rt_block_unmap(shim.ctx, c.block, c.layer);
  } else {
    if (m && !publish_locked(m))
      return false;
    rc = stg_rt_block_map(shim.ctx, c.block, c.model, c.layer, c.zmin, c.zmax, (int)(c.xy.size() / 2), c.xy.empty() ? nullptr : &c.xy[0]);
  }
  if (rc) {
    give_up(c.unmap ? "stg_rt_block_unmap" : "stg_rt_block_map", rc);
    return false;
  }
  return true;
}

// mu held. The context is created at the first sensor update: by then World::Load has rendered every model once, so
// the outlines seen so far give the extent (plus one SuperRegion, 1024 cells = 20 m at 2 cm, of room on every side).
bool ensure_context_locked(World *w)
{
  if (shim.ctx)
    return true;
  if (!shim.active())
    return false;
  if (shim.emulate) {
    shim.early.clear();
    shim.early_open = false;
    return true;
  }
  if (getenv("STAGE_GPU_FORCE_FALLBACK")) { // behave up to here like a run that will use the GPU, then do not: the
    shim.failed = true;                     // reference with the very heap layout (hence pointer order) of that run
    shim.early.clear();
    return false;
  }
  int32_t x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  bool any = false;
  for (size_t i = 0; i < shim.early.size(); i++)
    for (size_t k = 0; k + 1 < shim.early[i].xy.size(); k += 2) {
      const int32_t x = shim.early[i].xy[k], y = shim.early[i].xy[k + 1];
      if (!any) {
        x0 = x1 = x;
        y0 = y1 = y;
        any = true;
      }
      x0 = std::min(x0, x);
      x1 = std::max(x1, x);
      y0 = std::min(y0, y);
      y1 = std::max(y1, y);
    }
  stg_rt_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = sizeof(cfg);
  cfg.device = getenv("STAGE_GPU_DEVICE") ? atoi(getenv("STAGE_GPU_DEVICE")) : 0;
  cfg.ppm = w->ppm;
  cfg.sreg_x0 = (x0 >> 10) - 1;
  cfg.sreg_y0 = (y0 >> 10) - 1;
  cfg.sreg_nx = (x1 >> 10) - (x0 >> 10) + 3;
  cfg.sreg_ny = (y1 >> 10) - (y0 >> 10) + 3;
  uint32_t max_id = 0;
  for (std::set<Model *>::const_iterator it = w->models.begin(); it != w->models.end(); ++it)
    max_id = std::max(max_id, (*it)->id);
  cfg.max_models = 2 * (max_id + 1) + 1024;
  cfg.max_blocks = 2 * (uint32_t)shim.block_ids.size() + 4096;
  cfg.max_cell_entries = 0;
  cfg.flags = STG_RT_CFG_HOST_EXACT_FIDUCIALS; // controllers steer by bearing and range: the reference's own hypot / atan2
  const int rc = stg_rt_create(&cfg, &shim.ctx);
  if (rc) {
    shim.ctx = nullptr;
    give_up("stg_rt_create", rc);
    return false;
  }
  for (std::set<Model *>::const_iterator it = w->models.begin(); it != w->models.end(); ++it)
    if (!publish_locked(*it))
      return false;
  shim.published_at_update = w->updates;
  for (size_t i = 0; i < shim.early.size(); i++)
    if (!forward_map_locked(shim.early[i], nullptr))
      return false;
  if (shim.verbose)
    fprintf(stderr, "libstage_gpu: context on device %d, superregions [%d,%d)+[%d,%d), %zu early map calls replayed\n", cfg.device, cfg.sreg_x0,
            cfg.sreg_x0 + cfg.sreg_nx, cfg.sreg_y0, cfg.sreg_y0 + cfg.sreg_ny, shim.early.size());
  shim.early.clear();
  shim.early.shrink_to_fit();
  return true;
}

typedef void (*sensor_update_fn)(ModelRanger::Sensor *, ModelRanger *);
sensor_update_fn real_sensor_update()
{
  static sensor_update_fn fn = next_symbol<sensor_update_fn>("_ZN3Stg11ModelRanger6Sensor6UpdateEPS0_");
  return fn;
}

typedef void (*blob_update_fn)(ModelBlobfinder *);
blob_update_fn real_blob_update()
{
  static blob_update_fn fn = next_symbol<blob_update_fn>("_ZN3Stg15ModelBlobfinder6UpdateEv");
  return fn;
}

// The tick's GPU work is waited for and handed to the sensors; on any error the same sensors are updated by libstage's
// own loops instead (the layer they read, (updates + 1) % 2, is not written during the tick, so the answer is the same).
void finish_tick(World *w)
{
  std::unique_lock<std::mutex> lk(shim.mu);
  if (shim.rangers.empty() && shim.fid.empty() && shim.blobs.empty())
    return;
  std::vector<RangerJob> rangers;
  std::vector<FidBatch> fid;
  std::vector<BlobJob> blobs;
  rangers.swap(shim.rangers);
  fid.swap(shim.fid);
  blobs.swap(shim.blobs);
  if (shim.emulate) { // hand over what was held back, as the GPU path does after its sync
    shim.ticks++;
    lk.unlock();
    for (size_t i = 0; i < rangers.size(); i++) {
      ModelRanger::Sensor &s = *rangers[i].sensor;
      s.ranges = *rangers[i].gpu_ranges;
      s.intensities = *rangers[i].gpu_intensities;
      delete rangers[i].gpu_ranges;
      delete rangers[i].gpu_intensities;
      uint32_t in_range = 0;
      for (size_t t = 0; t < s.ranges.size(); t++)
        in_range += s.ranges[t] < s.range.max ? 1u : 0u;
      stg_rt_ranger_rand_advance((uint32_t)s.ranges.size(), in_range);
    }
    for (size_t b = 0; b < fid.size(); b++)
      for (size_t i = 0; i < fid[b].owners.size(); i++)
        fid[b].owners[i]->fiducials = fid[b].expected[i];
    for (size_t i = 0; i < blobs.size(); i++)
      blobs[i].owner->blobs = blobs[i].expected;
    return;
  }
  bool ok = shim.ctx != nullptr;
  for (size_t b = 0; b < fid.size() && ok; b++) {
    FidBatch &B = fid[b];
    B.max_per_finder = std::max<uint32_t>(1u, (uint32_t)std::min<size_t>(B.targets.size(), 4096));
    B.out.resize(B.finders.size() * (size_t)B.max_per_finder);
    B.count.assign(B.finders.size(), 0);
    const int rc = stg_rt_fiducials_submit(shim.ctx, (uint32_t)B.targets.size(), B.targets.empty() ? nullptr : &B.targets[0], (uint32_t)B.finders.size(),
                                           &B.finders[0], B.max_per_finder, &B.out[0], &B.count[0]);
    if (rc) {
      give_up("stg_rt_fiducials_submit", rc);
      ok = false;
    }
  }
  // all blobfinders of the tick in one batch; a finder's tracked colours are its stretch of one colour table
  std::vector<stg_rt_blob_finder> blob_recs(blobs.size());
  std::vector<double> blob_colors;
  std::vector<stg_rt_blob> blob_out;
  std::vector<uint32_t> blob_count(blobs.size(), 0);
  uint32_t blob_max = 1;
  if (ok && !blobs.empty()) {
    for (size_t i = 0; i < blobs.size(); i++) {
      blob_recs[i] = blobs[i].rec;
      blob_recs[i].first_color = (uint32_t)(blob_colors.size() / 3);
      blob_recs[i].n_colors = (uint32_t)(blobs[i].colors.size() / 3);
      blob_colors.insert(blob_colors.end(), blobs[i].colors.begin(), blobs[i].colors.end());
      blob_max = std::max(blob_max, blobs[i].rec.scan_width); // a blob needs at least one sample
    }
    blob_out.resize(blobs.size() * (size_t)blob_max);
    const int rc = stg_rt_blobs_submit(shim.ctx, (uint32_t)blobs.size(), &blob_recs[0], (uint32_t)(blob_colors.size() / 3),
                                       blob_colors.empty() ? nullptr : &blob_colors[0], blob_max, &blob_out[0], &blob_count[0]);
    if (rc) {
      give_up("stg_rt_blobs_submit", rc);
      ok = false;
    }
  }
  if (ok) {
    const int rc = stg_rt_sync(shim.ctx);
    if (rc) {
      give_up("stg_rt_sync", rc);
      ok = false;
    }
  }
  shim.ticks++;
  lk.unlock();
  // a blob as ModelBlobfinder::Update makes it (model_blobfinder.cc:231-238) from a record of the batch
  auto blob_of = [&](size_t i, uint32_t k) {
    const stg_rt_blob &r = blob_out[i * (size_t)blob_max + k];
    ModelBlobfinder::Blob b;
    const Model *seen = Model::LookupId(r.model);
    b.color = seen ? seen->GetColor() : Color();
    b.left = r.left;
    b.top = r.top;
    b.right = r.right;
    b.bottom = r.bottom;
    b.range = r.range;
    return b;
  };
  if (!ok && shim.verify) {
    for (size_t i = 0; i < rangers.size(); i++) {
      delete rangers[i].gpu_ranges;
      delete rangers[i].gpu_intensities;
    }
    return;
  }
  if (!ok) { // redo the tick's sensors on the CPU
    for (size_t i = 0; i < rangers.size(); i++)
      real_sensor_update()(rangers[i].sensor, rangers[i].mod);
    for (size_t b = 0; b < fid.size(); b++)
      for (size_t i = 0; i < fid[b].owners.size(); i++) {
        ModelFiducial *f = fid[b].owners[i];
        f->fiducials.clear();
        for (std::vector<Model *>::const_iterator it = w->models_with_fiducials.begin(); it != w->models_with_fiducials.end(); ++it)
          f->AddModelIfVisible(*it); // model_fiducial.cc:286-287, the reference's own alternative to the windowed search
      }
    for (size_t i = 0; i < blobs.size(); i++)
      real_blob_update()(blobs[i].owner); // (its Model::Update() sets last_update to the same sim_time once more)
    return;
  }
  if (shim.verify) { // the sensors hold libstage's own answers; the GPU's must be the same
    uint64_t beams = 0, bad_beams = 0, fids = 0, bad_updates = 0;
    double err = 0;
    for (size_t i = 0; i < rangers.size(); i++) {
      const ModelRanger::Sensor &s = *rangers[i].sensor;
      const std::vector<double> &gr = *rangers[i].gpu_ranges, &gi = *rangers[i].gpu_intensities;
      for (size_t t = 0; t < gr.size(); t++) {
        beams++;
        if (t >= s.ranges.size() || memcmp(&gr[t], &s.ranges[t], 8) || memcmp(&gi[t], &s.intensities[t], 8)) {
          if (bad_beams + shim.mismatched_beams < 8)
            fprintf(stderr, "libstage_gpu verify: update %llu model %u (%s) beam %zu of %zu: libstage %.17g / %.17g, gpu %.17g / %.17g\n",
                    (unsigned long long)w->updates, rangers[i].mod->id, rangers[i].mod->Token(), t, gr.size(),
                    t < s.ranges.size() ? s.ranges[t] : -1.0, t < s.intensities.size() ? s.intensities[t] : -1.0, gr[t], gi[t]);
          bad_beams++;
        }
      }
      delete rangers[i].gpu_ranges;
      delete rangers[i].gpu_intensities;
    }
    for (size_t b = 0; b < fid.size(); b++) {
      const FidBatch &B = fid[b];
      for (size_t i = 0; i < B.owners.size(); i++) {
        const std::vector<ModelFiducial::Fiducial> &want = B.expected[i];
        const uint32_t n = std::min(B.count[i], B.max_per_finder);
        bool same = n == want.size();
        for (uint32_t k = 0; k < n && same; k++) {
          const stg_rt_fiducial &r = B.out[i * (size_t)B.max_per_finder + k];
          const ModelFiducial::Fiducial &w = want[k];
          same = B.target_models[r.target] == w.mod && r.id == w.id && r.geom[0] == w.geom.x && r.geom[1] == w.geom.y && r.geom[2] == w.geom.z &&
                 r.geom[3] == w.geom.a && r.pose[0] == w.pose.x && r.pose[1] == w.pose.y && r.pose[2] == w.pose.z && r.pose[3] == w.pose.a &&
                 r.range == w.range && r.bearing == w.bearing; // STG_RT_CFG_HOST_EXACT_FIDUCIALS: the host's hypot / atan2
          err = std::max(err, std::max(fabs(r.range - w.range), fabs(r.bearing - w.bearing)));
        }
        fids += n;
        bad_updates += same ? 0 : 1;
      }
    }
    uint64_t n_blobs = 0, bad_blob_updates = 0;
    for (size_t i = 0; i < blobs.size(); i++) {
      const std::vector<ModelBlobfinder::Blob> &want = blobs[i].expected;
      const uint32_t n = std::min(blob_count[i], blob_max);
      bool same = n == want.size();
      for (uint32_t k = 0; k < n && same; k++) {
        const ModelBlobfinder::Blob got = blob_of(i, k);
        same = got.left == want[k].left && got.right == want[k].right && got.top == want[k].top && got.bottom == want[k].bottom &&
               !memcmp(&got.range, &want[k].range, 8) && got.color == want[k].color;
      }
      n_blobs += n;
      bad_blob_updates += same ? 0 : 1;
    }
    std::lock_guard<std::mutex> lk2(shim.mu);
    shim.verified_blobs += n_blobs;
    shim.mismatched_blob_updates += bad_blob_updates;
    shim.verified_beams += beams;
    shim.mismatched_beams += bad_beams;
    shim.verified_fiducials += fids;
    shim.mismatched_fiducial_updates += bad_updates;
    shim.fiducial_max_err = std::max(shim.fiducial_max_err, err);
    return;
  }
  // the libc rand() draws the beam loops would have made (include/stg_rt.h, stg_rt_ranger_rand_advance)
  for (size_t i = 0; i < rangers.size(); i++) {
    const ModelRanger::Sensor &s = *rangers[i].sensor;
    uint32_t in_range = 0;
    for (size_t t = 0; t < s.ranges.size(); t++)
      in_range += s.ranges[t] < s.range.max ? 1u : 0u;
    stg_rt_ranger_rand_advance((uint32_t)s.ranges.size(), in_range);
  }
  for (size_t b = 0; b < fid.size(); b++) {
    const FidBatch &B = fid[b];
    for (size_t i = 0; i < B.owners.size(); i++) {
      std::vector<ModelFiducial::Fiducial> &dst = B.owners[i]->fiducials;
      dst.clear();
      const uint32_t n = std::min(B.count[i], B.max_per_finder);
      for (uint32_t k = 0; k < n; k++) {
        const stg_rt_fiducial &r = B.out[i * (size_t)B.max_per_finder + k];
        ModelFiducial::Fiducial f;
        f.range = r.range;
        f.bearing = r.bearing;
        f.geom = Pose(r.geom[0], r.geom[1], r.geom[2], r.geom[3]);
        f.pose = Pose(r.pose[0], r.pose[1], r.pose[2], r.pose[3]);
        f.mod = B.target_models[r.target];
        f.id = r.id;
        dst.push_back(f);
      }
    }
  }
  for (size_t i = 0; i < blobs.size(); i++) {
    std::vector<ModelBlobfinder::Blob> &dst = blobs[i].owner->blobs;
    dst.clear();
    for (uint32_t k = 0, n = std::min(blob_count[i], blob_max); k < n; k++)
      dst.push_back(blob_of(i, k));
  }
}

} // namespace

// ================================================================================================
// the interposed members
// ================================================================================================

void Block::Map(unsigned int layer)
{
  typedef void (*fn_t)(Block *, unsigned int);
  static fn_t fn = next_symbol<fn_t>("_ZN3Stg5Block3MapEj");
  fn(this, layer); // the CPU grid and global_z, as ever
  if (!shim.active())
    return;
  const std::vector<point_int_t> px = group->mod.LocalToPixels(pts); // what World::MapPoly was just given
  static_assert(sizeof(point_int_t) == 2 * sizeof(int32_t), "point_int_t is two int32");
  MapCall c;
  c.unmap = false;
  c.model = group->mod.id;
  c.layer = (int)layer;
  c.zmin = global_z.min;
  c.zmax = global_z.max;
  c.xy.resize(2 * px.size());
  for (size_t i = 0; i < px.size(); i++) {
    c.xy[2 * i] = px[i].x;
    c.xy[2 * i + 1] = px[i].y;
  }
  std::lock_guard<std::mutex> lk(shim.mu);
  if (!shim.active())
    return;
  c.block = block_id_locked(this);
  shim.maps++;
  if (shim.emulate && !shim.early_open)
    return;
  if (!shim.ctx)
    shim.early.push_back(c);
  else
    forward_map_locked(c, &group->mod);
}

void Block::UnMap(unsigned int layer)
{
  typedef void (*fn_t)(Block *, unsigned int);
  static fn_t fn = next_symbol<fn_t>("_ZN3Stg5Block5UnMapEj");
  fn(this, layer);
  if (!shim.active())
    return;
  std::lock_guard<std::mutex> lk(shim.mu);
  if (!shim.active())
    return;
  MapCall c;
  c.unmap = true;
  c.block = block_id_locked(this);
  c.model = 0;
  c.layer = (int)layer;
  c.zmin = c.zmax = 0;
  if (shim.emulate && !shim.early_open)
    return;
  if (!shim.ctx)
    shim.early.push_back(c);
  else
    forward_map_locked(c, nullptr);
}

void ModelRanger::Sensor::Update(ModelRanger *mod)
{
  // a noisy sensor keeps libstage's loop: its exact libc rand() sequence cannot be replayed in parallel
  // (stg_rt_fans_submit_noisy gives the same distributions where that is enough - DESIGN.md section 4)
  bool gpu = shim.active() && angle_noise == 0.0 && range_noise == 0.0 && range_noise_const == 0.0 && sample_count > 0;
  if (gpu) {
    std::lock_guard<std::mutex> lk(shim.mu);
    gpu = ensure_context_locked(mod->world) && publish_changed_locked(mod->world);
  }
  if (!gpu) {
    real_sensor_update()(this, mod);
    return;
  }
  // the set-up of model_ranger.cc:213-230, then one fan record instead of the beam loop of :232-255
  ranges.resize(sample_count);
  intensities.resize(sample_count);
  bearings.resize(sample_count);
  const double sample_incr(fov / std::max(sample_count - 1, (unsigned int)1));
  const double start_angle = (sample_count > 1 ? -fov / 2.0 : 0.0);
  Pose rayorg(pose);
  rayorg.a += start_angle;
  rayorg.z += size.z / 2.0;
  rayorg = mod->LocalToGlobal(rayorg);
  for (size_t t(0); t < sample_count; t++) // :251 - a function of the sensor's geometry alone
    bearings[t] = start_angle + ((double)t) * sample_incr;
  stg_rt_fan f;
  memset(&f, 0, sizeof(f));
  f.finder = mod->id;
  f.n_samples = sample_count;
  f.pred = STG_RT_PRED_RANGER;
  f.ztest = 1;
  f.layer = (uint8_t)((mod->world->updates + 1) % 2); // world.cc:804
  f.angles = STG_RT_ANGLES_RANGER;
  f.ox = rayorg.x;
  f.oy = rayorg.y;
  f.oz = rayorg.z;
  f.oa = rayorg.a;
  f.range = range.max;
  f.fov = fov;
  stg_rt_fan_out o;
  memset(&o, 0, sizeof(o));
  o.ranges = &ranges[0];
  o.intensities = &intensities[0];
  RangerJob j;
  j.sensor = this;
  j.mod = mod;
  j.gpu_ranges = j.gpu_intensities = nullptr;
  if (shim.emulate) { // libstage's own loop, its answer held back and its rand() draws undone
    RngSnapshot snap;
    rng_save(snap);
    real_sensor_update()(this, mod);
    rng_restore(snap);
    j.gpu_ranges = new std::vector<double>(ranges);
    j.gpu_intensities = new std::vector<double>(intensities);
    std::fill(ranges.begin(), ranges.end(), -1.0); // nobody may look before the sync point
    std::fill(intensities.begin(), intensities.end(), -1.0);
    std::lock_guard<std::mutex> lk(shim.mu);
    shim.rangers.push_back(j);
    shim.fans++;
    return;
  }
  if (shim.verify) { // libstage's own loop fills the sensor now (and makes its rand() draws); the GPU answers on the side
    real_sensor_update()(this, mod);
    j.gpu_ranges = new std::vector<double>(sample_count);
    j.gpu_intensities = new std::vector<double>(sample_count);
    o.ranges = &(*j.gpu_ranges)[0];
    o.intensities = &(*j.gpu_intensities)[0];
  }
  std::unique_lock<std::mutex> lk(shim.mu);
  int rc = shim.ctx ? stg_rt_fans_submit(shim.ctx, 1, &f, nullptr, nullptr, &o) : STG_RT_ESTATE;
  if (rc) {
    give_up("stg_rt_fans_submit", rc);
    lk.unlock();
    if (!shim.verify)
      real_sensor_update()(this, mod);
    delete j.gpu_ranges;
    delete j.gpu_intensities;
    return;
  }
  shim.rangers.push_back(j);
  shim.fans++;
}

void ModelFiducial::Update(void)
{
  typedef void (*fn_t)(ModelFiducial *);
  static fn_t fn = next_symbol<fn_t>("_ZN3Stg13ModelFiducial6UpdateEv");
  bool gpu = shim.active();
  if (gpu) {
    std::lock_guard<std::mutex> lk(shim.mu);
    gpu = ensure_context_locked(world) && publish_changed_locked(world);
  }
  if (!gpu) {
    fn(this);
    return;
  }
  if (subs < 1) // model_fiducial.cc:233-234
    return;
  std::vector<Fiducial> expected;
  if (shim.verify || shim.emulate) { // libstage's own Update, Model::Update() included; its records are what the GPU's must equal
    fn(this);
    expected = fiducials;
    if (shim.emulate)
      fiducials.clear(); // held back until the sync point
  } else {
    fiducials.clear(); // :237; filled at the sync point
  }
  // what AddModelIfVisible reads (:109-224), taken NOW: the finder's pose, and every fiducial-bearing model of the
  // world in ascending pointer order - the order the reference visits candidates in (:266-283)
  const Pose my = GetGlobalPose();
  const Pose org = GetGlobalPose() + geom.pose; // the frame Model::LocalToGlobal composes the ray pose on
  stg_rt_fid_finder F;
  memset(&F, 0, sizeof(F));
  F.finder = id;
  F.fiducial_key = vis.fiducial_key;
  F.ignore_zloc = ignore_zloc ? 1 : 0;
  F.layer = (uint8_t)((world->updates + 1) % 2);
  F.x = my.x; F.y = my.y; F.z = my.z; F.a = my.a;
  F.ox = org.x; F.oy = org.y; F.oz = org.z; F.oa = org.a;
  F.max_range_anon = max_range_anon;
  F.max_range_id = max_range_id;
  F.fov = fov;
  std::vector<Model *> models(world->models_with_fiducials.begin(), world->models_with_fiducials.end());
  std::sort(models.begin(), models.end());
  std::vector<stg_rt_fid_target> targets(models.size());
  for (size_t i = 0; i < models.size(); i++) {
    stg_rt_fid_target &t = targets[i];
    memset(&t, 0, sizeof(t));
    const Pose hp = models[i]->GetGlobalPose();
    const Geom hg = models[i]->GetGeom();
    t.model = models[i]->id;
    t.fiducial_return = models[i]->vis.fiducial_return;
    t.fiducial_key = models[i]->vis.fiducial_key;
    t.x = hp.x; t.y = hp.y; t.z = hp.z; t.a = hp.a;
    t.sx = hg.size.x; t.sy = hg.size.y; t.sz = hg.size.z;
  }
  {
    std::lock_guard<std::mutex> lk(shim.mu);
    FidBatch *B = shim.fid.empty() ? nullptr : &shim.fid.back();
    if (!B || B->targets.size() != targets.size() ||
        (!targets.empty() && memcmp(&B->targets[0], &targets[0], targets.size() * sizeof(stg_rt_fid_target)))) {
      shim.fid.push_back(FidBatch());
      B = &shim.fid.back();
      B->targets.swap(targets);
      B->target_models.swap(models);
    }
    B->finders.push_back(F);
    B->owners.push_back(this);
    B->expected.push_back(expected);
    shim.fid_updates++;
  }
  if (!shim.verify && !shim.emulate)
    Model::Update(); // :292
}

void ModelBlobfinder::Update(void)
{
  bool gpu = shim.active();
  if (gpu) {
    std::lock_guard<std::mutex> lk(shim.mu);
    gpu = ensure_context_locked(world) && publish_changed_locked(world);
  }
  if (!gpu) {
    real_blob_update()(this);
    return;
  }
  BlobJob j;
  j.owner = this;
  if (shim.verify || shim.emulate) { // libstage's own Update, Model::Update() included
    real_blob_update()(this);
    j.expected = blobs;
    if (shim.emulate)
      blobs.clear(); // held back until the sync point
  } else {
    blobs.clear(); // model_blobfinder.cc:175; filled at the sync point
  }
  // Raytrace(Pose(0, 0, 0, pan), range, fov, blob_match, NULL, false, samples), :170 - the pose the fan leaves from, taken NOW
  const Pose origin = LocalToGlobal(Pose(0, 0, 0, pan));
  memset(&j.rec, 0, sizeof(j.rec));
  j.rec.finder = id;
  j.rec.scan_width = scan_width;
  j.rec.scan_height = scan_height;
  j.rec.layer = (uint8_t)((world->updates + 1) % 2);
  j.rec.ox = origin.x; j.rec.oy = origin.y; j.rec.oz = origin.z; j.rec.oa = origin.a;
  j.rec.range = range;
  j.rec.fov = fov;
  for (size_t c = 0; c < colors.size(); c++) {
    j.colors.push_back(colors[c].r);
    j.colors.push_back(colors[c].g);
    j.colors.push_back(colors[c].b);
  }
  {
    std::lock_guard<std::mutex> lk(shim.mu);
    shim.blobs.push_back(j);
    shim.blob_updates++;
  }
  if (!shim.verify && !shim.emulate)
    Model::Update(); // :249
}

void World::CallUpdateCallbacks()
{
  typedef void (*fn_t)(World *);
  static fn_t fn = next_symbol<fn_t>("_ZN3Stg5World19CallUpdateCallbacksEv");
  if (shim.wanted)
    finish_tick(this);
  fn(this);
}

// for the driver's summary line (integration/stage_run.cc): how much of the run went through the GPU
extern "C" void stage_gpu_counters(unsigned long long out[13])
{
  std::lock_guard<std::mutex> lk(shim.mu);
  out[0] = shim.fans;
  out[1] = shim.fid_updates;
  out[2] = shim.maps;
  out[3] = shim.ticks;
  out[4] = shim.failed ? 1 : (shim.emulate ? 4 : shim.ctx ? (shim.verify ? 3 : 2) : 0);
  out[5] = shim.verified_beams;
  out[6] = shim.mismatched_beams;
  out[7] = shim.verified_fiducials;
  out[8] = shim.mismatched_fiducial_updates;
  out[9] = (unsigned long long)(shim.fiducial_max_err * 1e18); // in attometres / attoradians: fits an integer
  out[10] = shim.blob_updates;
  out[11] = shim.verified_blobs;
  out[12] = shim.mismatched_blob_updates;
}
