// T0 oracle harness: drives the UNMODIFIED reference (oracle/_ref/libstage_ref.so, built by
// oracle/build_ref.sh from /root/reference) and records what its raycast hot path does.
// TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs through ctypes. Never linked into stage_b200.
//
// How the recording works without touching reference sources: libstage_ref.so is built -fPIC with
// default (interposable) symbol visibility, so its internal calls to Block::Map, Block::UnMap and
// World::Raytrace(const Ray&) go through the PLT. This library defines those three symbols
// itself; being earlier in the lookup scope it is bound first, logs the call, and forwards to the
// real definition found with dlsym() on the libstage_ref.so handle.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <dlfcn.h>

// the harness needs Block::pts, Block::global_z, World::models ... (all private/protected)
#define private public
#define protected public
#include "stage.hh"
#include "region.hh"
#undef private
#undef protected

#include "trace_format.h"

using namespace Stg;

namespace {

struct Recorder {
  std::mutex mu;
  bool active = false;
  std::vector<TraceEvent> events;
  std::vector<int32_t> verts;
  std::vector<TraceRay> rays;
  std::vector<double> aux;
  std::map<const Block *, uint32_t> block_ids;
  std::map<ray_test_func_t, int> kinds;
  uint32_t n_ticks = 0;
  World *world = NULL;
} rec;

thread_local int tl_quiet = 0; // >0: calls made by the harness itself are not recorded
bool g_fiducials_on_main = false;

void *ref_handle()
{
  static void *h = dlopen("libstage_ref.so", RTLD_NOW | RTLD_NOLOAD);
  if (!h) {
    fprintf(stderr, "ref_harness: libstage_ref.so not loaded: %s\n", dlerror());
    abort();
  }
  return h;
}

template <class F> F real(const char *mangled)
{
  void *p = dlsym(ref_handle(), mangled);
  if (!p) {
    fprintf(stderr, "ref_harness: missing %s\n", mangled);
    abort();
  }
  return (F)p;
}

uint32_t block_id_locked(const Block *b)
{
  std::map<const Block *, uint32_t>::iterator it = rec.block_ids.find(b);
  if (it != rec.block_ids.end())
    return it->second;
  uint32_t id = (uint32_t)rec.block_ids.size();
  rec.block_ids[b] = id;
  return id;
}

// harness-owned predicates with the same meaning as the reference's file-static ones
bool h_ranger_match(Model *hit, const Model *finder, const void *)
{ // model_ranger.cc:158-170
  if ((hit == finder->Parent()) || (hit == finder))
    return false;
  return ((!hit->IsRelated(finder)) && (sgn(hit->vis.ranger_return) != -1));
}
bool h_unrelated_match(Model *hit, const Model *finder, const void *)
{ // model_fiducial.cc:103-107
  return (!finder->IsRelated(hit));
}
bool h_gripper_match(Model *hit, const Model *finder, const void *)
{ // model_gripper.cc:329-336
  return ((hit != finder) && hit->vis.gripper_return);
}

int classify(const Ray &r)
{
  if (r.func == h_ranger_match)
    return TRACE_KIND_RANGER;
  if (r.func == h_unrelated_match)
    return TRACE_KIND_UNRELATED;
  if (r.func == h_gripper_match)
    return TRACE_KIND_GRIPPER;
  std::map<ray_test_func_t, int>::iterator it = rec.kinds.find(r.func);
  if (it != rec.kinds.end())
    return it->second;
  // the reference's predicates are file-static; tell them apart by who is looking
  const std::string &t = r.mod->type;
  int k = (t == "ranger") ? TRACE_KIND_RANGER : (t == "gripper") ? TRACE_KIND_GRIPPER : TRACE_KIND_UNRELATED;
  rec.kinds[r.func] = k;
  return k;
}

uint32_t root_id(const Model *m)
{
  while (m->parent)
    m = m->parent;
  return m->id;
}

} // namespace

// ---------------------------------------------------------------- interposed reference symbols

void Block::Map(unsigned int layer)
{
  typedef void (*fn_t)(Block *, unsigned int);
  static fn_t fn = real<fn_t>("_ZN3Stg5Block3MapEj");
  fn(this, layer);
  if (!rec.active || tl_quiet)
    return;
  // what MapPoly was given (block.cc:174) and the z bounds Map left behind (block.cc:177-180)
  const std::vector<point_int_t> px = group->mod.LocalToPixels(pts);
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_MAP;
  e.layer = (uint8_t)layer;
  e.n_pts = (uint16_t)px.size();
  e.block = block_id_locked(this);
  e.model = group->mod.id;
  e.first = (uint32_t)(rec.verts.size() / 2);
  e.zmin = global_z.min;
  e.zmax = global_z.max;
  for (size_t i = 0; i < px.size(); i++) {
    rec.verts.push_back(px[i].x);
    rec.verts.push_back(px[i].y);
  }
  rec.events.push_back(e);
}

void Block::UnMap(unsigned int layer)
{
  typedef void (*fn_t)(Block *, unsigned int);
  static fn_t fn = real<fn_t>("_ZN3Stg5Block5UnMapEj");
  const bool was_mapped = !rendered_cells[layer].empty();
  fn(this, layer);
  if (!rec.active || tl_quiet || !was_mapped)
    return;
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_UNMAP;
  e.layer = (uint8_t)layer;
  e.block = block_id_locked(this);
  e.model = group->mod.id;
  rec.events.push_back(e);
}

static thread_local int tl_collision_depth = 0;

Model *Model::TestCollision()
{
  typedef Model *(*fn_t)(Model *);
  static fn_t fn = real<fn_t>("_ZN3Stg5Model13TestCollisionEv");
  tl_collision_depth++; // the children's tests come through here too (model.cc:672): log the top call
  Model *hit = fn(this);
  tl_collision_depth--;
  if (tl_collision_depth || !rec.active || tl_quiet)
    return hit;
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_COLLISION;
  e.layer = (uint8_t)(world->updates % 2);
  e.model = id;
  e.block = hit ? hit->id + 1 : 0;
  e.first = (uint32_t)rec.aux.size();
  std::vector<Model *> todo(1, this); // depth first, children in order
  uint32_t n = 0;
  while (!todo.empty()) {
    Model *m = todo.back();
    todo.pop_back();
    for (size_t i = 0; i < m->blockgroup.blocks.size(); i++, n++)
      rec.aux.push_back((double)block_id_locked(&m->blockgroup.blocks[i]));
    for (size_t i = m->children.size(); i-- > 0;)
      todo.push_back(m->children[i]);
  }
  e.n_pts = (uint16_t)n;
  rec.events.push_back(e);
  return hit;
}

void Model::AppendTouchingModels(std::set<Model *> &touchers)
{
  typedef void (*fn_t)(Model *, std::set<Model *> &);
  static fn_t fn = real<fn_t>("_ZN3Stg5Model20AppendTouchingModelsERSt3setIPS0_St4lessIS2_ESaIS2_EE");
  std::set<Model *> mine; // what THIS call adds, whatever the caller's set already held
  fn(this, mine);
  touchers.insert(mine.begin(), mine.end());
  if (!rec.active || tl_quiet)
    return;
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_TOUCHERS;
  e.layer = (uint8_t)(world->updates % 2);
  e.model = id;
  e.first = (uint32_t)rec.aux.size();
  e.n_pts = (uint16_t)blockgroup.blocks.size();
  for (size_t i = 0; i < blockgroup.blocks.size(); i++)
    rec.aux.push_back((double)block_id_locked(&blockgroup.blocks[i]));
  std::vector<uint32_t> ids;
  for (std::set<Model *>::iterator it = mine.begin(); it != mine.end(); ++it)
    ids.push_back((*it)->id);
  std::sort(ids.begin(), ids.end());
  e.block = (uint32_t)ids.size();
  for (size_t i = 0; i < ids.size(); i++)
    rec.aux.push_back((double)ids[i]);
  rec.events.push_back(e);
}

RaytraceResult World::Raytrace(const Ray &r)
{
  typedef RaytraceResult (*fn_t)(World *, const Ray &);
  static fn_t fn = real<fn_t>("_ZN3Stg5World8RaytraceERKNS_3RayE");
  const unsigned int layer = (updates + 1) % 2; // world.cc:804
  RaytraceResult res = fn(this, r);
  if (!rec.active || tl_quiet)
    return res;
  std::lock_guard<std::mutex> g(rec.mu);
  TraceRay t;
  memset(&t, 0, sizeof(t));
  t.ox = r.origin.x;
  t.oy = r.origin.y;
  t.oz = r.origin.z;
  t.oa = r.origin.a;
  t.range = r.range;
  t.res_range = res.range;
  t.hit_model = res.mod ? (int32_t)res.mod->id : -1;
  t.finder = r.mod->id;
  t.kind = (uint8_t)classify(r);
  t.ztest = r.ztest ? 1 : 0;
  t.layer = (uint8_t)layer;
  if (!rec.events.empty() && rec.events.back().type == TRACE_EV_RAYS) {
    rec.events.back().model++;
  } else {
    TraceEvent e;
    memset(&e, 0, sizeof(e));
    e.type = TRACE_EV_RAYS;
    e.model = 1;
    e.first = (uint32_t)rec.rays.size();
    rec.events.push_back(e);
  }
  rec.rays.push_back(t);
  return res;
}

void ModelRanger::Sensor::Update(ModelRanger *mod)
{
  typedef void (*fn_t)(ModelRanger::Sensor *, ModelRanger *);
  static fn_t fn = real<fn_t>("_ZN3Stg11ModelRanger6Sensor6UpdateEPS0_");
  // what a patched libstage would hand to stg_rt_fans_submit: the pose of the first ray,
  // computed exactly as model_ranger.cc:222-230 does before the beam loop
  const double start_angle = (sample_count > 1 ? -fov / 2.0 : 0.0);
  Pose rayorg(pose);
  rayorg.a += start_angle;
  rayorg.z += size.z / 2.0;
  rayorg = mod->LocalToGlobal(rayorg);
  const unsigned int layer = (mod->world->updates + 1) % 2;
  fn(this, mod);
  if (!rec.active || tl_quiet || range_noise != 0 || range_noise_const != 0 || angle_noise != 0)
    return;
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_RANGER_SCAN;
  e.layer = (uint8_t)layer;
  e.n_pts = (uint16_t)sample_count;
  e.model = mod->id;
  for (size_t i = 0; i < mod->sensors.size(); i++)
    if (&mod->sensors[i] == this)
      e.block = (uint32_t)i;
  e.first = (uint32_t)rec.aux.size();
  const double hdr[TRACE_SCAN_HEADER_DOUBLES] = { rayorg.x, rayorg.y, rayorg.z, rayorg.a, range.max, fov };
  rec.aux.insert(rec.aux.end(), hdr, hdr + TRACE_SCAN_HEADER_DOUBLES);
  rec.aux.insert(rec.aux.end(), ranges.begin(), ranges.end());
  rec.aux.insert(rec.aux.end(), intensities.begin(), intensities.end());
  rec.aux.insert(rec.aux.end(), bearings.begin(), bearings.end());
  rec.events.push_back(e);
}

unsigned int World::GetEventQueue(Model *mod) const
{
  typedef unsigned int (*fn_t)(const World *, Model *);
  static fn_t fn = real<fn_t>("_ZNK3Stg5World13GetEventQueueEPNS_5ModelE");
  if (g_fiducials_on_main && dynamic_cast<ModelFiducial *>(mod))
    return 0;
  return fn(this, mod);
}

void ModelFiducial::Update(void)
{
  typedef void (*fn_t)(ModelFiducial *);
  static fn_t fn = real<fn_t>("_ZN3Stg13ModelFiducial6UpdateEv");
  if (!rec.active || tl_quiet || subs < 1) {
    fn(this);
    return;
  }
  // inputs, as a patched libstage would gather them for stg_rt_fiducials_submit
  const Pose my = GetGlobalPose();
  const Pose org = GetGlobalPose() + geom.pose; // the frame Model::LocalToGlobal composes on
  const unsigned int layer = (world->updates + 1) % 2;
  std::vector<Model *> targets(world->models_with_fiducials.begin(), world->models_with_fiducials.end());
  std::sort(targets.begin(), targets.end()); // candidates are visited in ascending Model* order
  std::vector<double> a;
  const double hdr[TRACE_FID_HEADER_DOUBLES] = { my.x, my.y, my.z, my.a, org.x, org.y, org.z, org.a, max_range_anon,
                                                 max_range_id, fov, (double)vis.fiducial_key,
                                                 ignore_zloc ? 1.0 : 0.0, (double)targets.size() };
  a.insert(a.end(), hdr, hdr + TRACE_FID_HEADER_DOUBLES);
  for (size_t i = 0; i < targets.size(); i++) {
    const Model *m = targets[i];
    const Pose p = m->GetGlobalPose();
    const Size sz = m->GetGeom().size;
    const double t[TRACE_FID_TARGET_DOUBLES] = { (double)m->id, (double)m->vis.fiducial_return, (double)m->vis.fiducial_key,
                                                 p.x, p.y, p.z, p.a, sz.x, sz.y, sz.z };
    a.insert(a.end(), t, t + TRACE_FID_TARGET_DOUBLES);
  }
  fn(this);
  for (size_t i = 0; i < fiducials.size(); i++) {
    const Fiducial &f = fiducials[i];
    const double r[TRACE_FID_RESULT_DOUBLES] = { (double)f.mod->id, (double)f.id, f.range, f.bearing, f.geom.x, f.geom.y,
                                                 f.geom.z, f.geom.a, f.pose.x, f.pose.y, f.pose.z, f.pose.a };
    a.insert(a.end(), r, r + TRACE_FID_RESULT_DOUBLES);
  }
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_FIDUCIAL;
  e.layer = (uint8_t)layer;
  e.model = id;
  e.block = (uint32_t)fiducials.size();
  e.first = (uint32_t)rec.aux.size();
  rec.aux.insert(rec.aux.end(), a.begin(), a.end());
  rec.events.push_back(e);
}

void ModelBlobfinder::Update(void)
{
  typedef void (*fn_t)(ModelBlobfinder *);
  static fn_t fn = real<fn_t>("_ZN3Stg15ModelBlobfinder6UpdateEv");
  if (!rec.active || tl_quiet) {
    fn(this);
    return;
  }
  const Pose org = LocalToGlobal(Pose(0, 0, 0, pan)); // what Model::Raytrace hands to World::Raytrace (stage.hh:2012)
  const unsigned int layer = (world->updates + 1) % 2;
  fn(this);
  std::lock_guard<std::mutex> g(rec.mu);
  TraceEvent e;
  memset(&e, 0, sizeof(e));
  e.type = TRACE_EV_BLOB;
  e.layer = (uint8_t)layer;
  e.model = id;
  e.block = (uint32_t)blobs.size();
  e.first = (uint32_t)rec.aux.size();
  const double hdr[TRACE_BLOB_HEADER_DOUBLES] = { org.x, org.y, org.z, org.a, range, fov, (double)scan_width,
                                                  (double)scan_height, (double)colors.size() };
  rec.aux.insert(rec.aux.end(), hdr, hdr + TRACE_BLOB_HEADER_DOUBLES);
  for (size_t i = 0; i < colors.size(); i++) {
    rec.aux.push_back(colors[i].r);
    rec.aux.push_back(colors[i].g);
    rec.aux.push_back(colors[i].b);
  }
  for (size_t i = 0; i < blobs.size(); i++) {
    const Blob &b = blobs[i];
    const double r[TRACE_BLOB_RESULT_DOUBLES] = { b.color.r, b.color.g, b.color.b, (double)b.left, (double)b.top,
                                                  (double)b.right, (double)b.bottom, b.range };
    rec.aux.insert(rec.aux.end(), r, r + TRACE_BLOB_RESULT_DOUBLES);
  }
  rec.events.push_back(e);
}

// ---------------------------------------------------------------- C interface for ctypes

extern "C" {

typedef struct {
  double ox, oy, oz, oa, range;
  uint32_t finder;
  uint8_t kind, ztest, pad[2];
} RefRayIn; // 48 B

typedef struct {
  double range;
  int32_t hit_model;
  int32_t pad;
} RefRayOut; // 16 B

static bool g_inited = false;
static void ensure_init()
{
  if (g_inited)
    return;
  int argc = 1;
  char arg0[] = "ref_harness";
  char *argv0[] = { arg0, NULL };
  char **argv = argv0;
  Stg::Init(&argc, &argv);
  g_inited = true;
}

// Start recording. Everything the reference maps / unmaps / traces from here on is logged.
void ref_trace_begin(void)
{
  std::lock_guard<std::mutex> g(rec.mu);
  rec.events.clear();
  rec.verts.clear();
  rec.rays.clear();
  rec.aux.clear();
  rec.block_ids.clear();
  rec.kinds.clear();
  rec.n_ticks = 0;
  rec.active = true;
}

// Load a world from a file (World::Load, world.cc:380-400). The World is leaked on purpose:
// ~World crashes in the reference itself.
void *ref_world_load_file(const char *path)
{
  ensure_init();
  World *w = new World("oracle");
  if (!w->Load(std::string(path)))
    return NULL;
  rec.world = w;
  return w;
}

// Load a world from worldfile text held in memory (World::Load(istream&), world.cc:385).
void *ref_world_load_text(const char *text, const char *path_hint)
{
  ensure_init();
  World *w = new World("oracle");
  std::istringstream ss(text);
  if (!w->Load(ss, std::string(path_hint ? path_hint : "")))
    return NULL;
  rec.world = w;
  return w;
}

// Run n simulation ticks (World::Update, world.cc:605-676), logging a TICK event after each.
// Returns the number of ticks actually run (stops early at quit_time).
int ref_world_update(void *wv, int n)
{
  World *w = (World *)wv;
  int done = 0;
  for (int i = 0; i < n; i++) {
    bool quit = w->Update();
    done++;
    if (rec.active) {
      std::lock_guard<std::mutex> g(rec.mu);
      TraceEvent e;
      memset(&e, 0, sizeof(e));
      e.type = TRACE_EV_TICK;
      e.block = (uint32_t)w->updates;
      rec.events.push_back(e);
      rec.n_ticks++;
    }
    if (quit)
      break;
  }
  return done;
}

double ref_world_ppm(void *wv) { return ((World *)wv)->ppm; }
uint64_t ref_world_updates(void *wv) { return ((World *)wv)->updates; }

// Stop recording and write the trace (oracle/trace_format.h). Returns 0 on success.
int ref_trace_end(void *wv, const char *path)
{
  World *w = (World *)wv;
  std::lock_guard<std::mutex> g(rec.mu);
  rec.active = false;
  std::vector<TraceModel> models;
  // World::models is a std::set<Model*> (stage.hh:788); emit sorted by id for stable files
  for (std::set<Model *>::iterator it = w->models.begin(); it != w->models.end(); ++it) {
    const Model *m = *it;
    TraceModel tm;
    memset(&tm, 0, sizeof(tm));
    tm.id = m->id;
    tm.parent = m->parent ? m->parent->id : 0xffffffffu;
    tm.root = root_id(m);
    tm.fiducial_return = m->vis.fiducial_return;
    tm.fiducial_key = m->vis.fiducial_key;
    tm.flags = (m->vis.gripper_return ? TRACE_FLAG_GRIPPER_RETURN : 0) |
               (m->vis.obstacle_return ? TRACE_FLAG_OBSTACLE_RETURN : 0) |
               (m->vis.blob_return ? TRACE_FLAG_BLOB_RETURN : 0);
    tm.ranger_return = m->vis.ranger_return;
    tm.color[0] = m->color.r;
    tm.color[1] = m->color.g;
    tm.color[2] = m->color.b;
    tm.color[3] = m->color.a;
    strncpy(tm.type, m->type.c_str(), sizeof(tm.type) - 1);
    models.push_back(tm);
  }
  std::sort(models.begin(), models.end(),
            [](const TraceModel &a, const TraceModel &b) { return a.id < b.id; });
  TraceHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, STG_TRACE_MAGIC, 8);
  h.ppm = w->ppm;
  h.n_models = (uint32_t)models.size();
  h.n_events = (uint32_t)rec.events.size();
  h.n_verts = (uint32_t)(rec.verts.size() / 2);
  h.n_rays = (uint32_t)rec.rays.size();
  h.n_ticks = rec.n_ticks;
  h.n_aux = (uint32_t)rec.aux.size();
  FILE *f = fopen(path, "wb");
  if (!f)
    return -1;
  fwrite(&h, sizeof(h), 1, f);
  fwrite(models.data(), sizeof(TraceModel), models.size(), f);
  fwrite(rec.events.data(), sizeof(TraceEvent), rec.events.size(), f);
  fwrite(rec.verts.data(), sizeof(int32_t), rec.verts.size(), f);
  fwrite(rec.rays.data(), sizeof(TraceRay), rec.rays.size(), f);
  fwrite(rec.aux.data(), sizeof(double), rec.aux.size(), f);
  fclose(f);
  return 0;
}

// Trace n rays with the reference's own World::Raytrace (world.cc:746-754 -> :756-963), split
// over nthreads host threads (the world must be quiescent: Raytrace is read-only). Not recorded.
// Returns the wall-clock seconds spent inside the tracing loops (max over threads).
double ref_raytrace_batch(void *wv, uint64_t n, const RefRayIn *in, RefRayOut *out, int nthreads)
{
  World *w = (World *)wv;
  if (nthreads < 1)
    nthreads = 1;
  std::vector<Model *> byid;
  for (std::set<Model *>::iterator it = w->models.begin(); it != w->models.end(); ++it) {
    if ((*it)->id >= byid.size())
      byid.resize((*it)->id + 1, NULL);
    byid[(*it)->id] = *it;
  }
  std::vector<double> secs(nthreads, 0.0);
  auto work = [&](int t) {
    tl_quiet++;
    const uint64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
    auto t0 = std::chrono::steady_clock::now();
    for (uint64_t i = lo; i < hi; i++) {
      const RefRayIn &r = in[i];
      ray_test_func_t fn = r.kind == TRACE_KIND_RANGER
                               ? h_ranger_match
                               : r.kind == TRACE_KIND_GRIPPER ? h_gripper_match : h_unrelated_match;
      RaytraceResult res =
          w->Raytrace(Pose(r.ox, r.oy, r.oz, r.oa), r.range, fn, byid[r.finder], NULL, r.ztest != 0);
      out[i].range = res.range;
      out[i].hit_model = res.mod ? (int32_t)res.mod->id : -1;
    }
    secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    tl_quiet--;
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
      th.emplace_back(work, t);
    for (auto &x : th)
      x.join();
  }
  return *std::max_element(secs.begin(), secs.end());
}

// Dump the reference's occupancy grid: one record per (layer, cell, position-in-list) entry of
// Cell::blocks[layer] (region.hh:47), in list order. Blocks are named by the recorder's dense ids
// (so call this between ref_trace_begin and ref_trace_end). out may be NULL to just count.
// Record = {layer, x, y, pos, block} as int32[5]. Also returns Region::count per region when
// counts != NULL: {rx, ry, count} int64[3] for regions with count > 0.
int64_t ref_dump_grid(void *wv, int32_t *out, int64_t cap, int64_t *counts, int64_t counts_cap,
                      int64_t *n_counts)
{
  World *w = (World *)wv;
  int64_t n = 0, nc = 0;
  std::lock_guard<std::mutex> g(rec.mu);
  for (std::map<point_int_t, SuperRegion *>::iterator it = w->superregions.begin();
       it != w->superregions.end(); ++it) {
    SuperRegion *sr = it->second;
    for (int ry = 0; ry < SUPERREGIONWIDTH; ry++)
      for (int rx = 0; rx < SUPERREGIONWIDTH; rx++) {
        Region *reg = sr->GetRegion(rx, ry);
        const int64_t gx = ((int64_t)sr->origin.x << SBITS) + rx, gy = ((int64_t)sr->origin.y << SBITS) + ry;
        if (reg->count) {
          if (counts && nc < counts_cap) {
            counts[3 * nc] = gx;
            counts[3 * nc + 1] = gy;
            counts[3 * nc + 2] = (int64_t)reg->count;
          }
          nc++;
        }
        if (reg->cells.empty())
          continue;
        for (int cy = 0; cy < REGIONWIDTH; cy++)
          for (int cx = 0; cx < REGIONWIDTH; cx++) {
            Cell &c = reg->cells[cx + cy * REGIONWIDTH];
            for (int layer = 0; layer < 2; layer++)
              for (size_t p = 0; p < c.blocks[layer].size(); p++) {
                if (out && n < cap) {
                  int32_t *o = out + 5 * n;
                  o[0] = layer;
                  o[1] = (int32_t)(gx * REGIONWIDTH + cx);
                  o[2] = (int32_t)(gy * REGIONWIDTH + cy);
                  o[3] = (int32_t)p;
                  std::map<const Block *, uint32_t>::iterator bi = rec.block_ids.find(c.blocks[layer][p]);
                  o[4] = bi == rec.block_ids.end() ? -1 : (int32_t)bi->second;
                }
                n++;
              }
          }
      }
  }
  if (n_counts)
    *n_counts = nc;
  return n;
}

// Model::Subscribe (model.cc:521): what a controller or a Player client does to switch a sensor on.
int ref_model_subscribe(void *wv, const char *name)
{
  World *w = (World *)wv;
  Model *m = w->GetModel(std::string(name));
  if (!m)
    return -1;
  m->Subscribe();
  return 0;
}

// What ModelPosition::Move does to the grid for one mover that hits nothing (model_position.cc:541-552): take the
// new pose, UnMapWithChildren(layer), MapWithChildren(layer) - for n models, one after the other on the calling
// thread, as World::Update runs its movers (world.cc:647-648). Returns the seconds it took. The models need not be
// position models (the benchmark's robots are plain models with a body). Built with -fno-access-control.
// record != 0: the Map / UnMap calls go into the trace like the reference's own (tests that replay them).
double ref_models_remap2(void *wv, uint32_t n, const uint32_t *ids, const double *xyza, int layer, int record)
{
  (void)wv;
  if (!record)
    tl_quiet++;
  auto t0 = std::chrono::steady_clock::now();
  for (uint32_t i = 0; i < n; i++) {
    Model *m = Model::LookupId(ids[i]);
    if (!m)
      continue;
    m->pose = Pose(xyza[4 * i], xyza[4 * i + 1], xyza[4 * i + 2], xyza[4 * i + 3]);
    m->UnMapWithChildren((unsigned int)layer);
    m->MapWithChildren((unsigned int)layer);
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!record)
    tl_quiet--;
  return secs;
}

double ref_models_remap(void *wv, uint32_t n, const uint32_t *ids, const double *xyza, int layer)
{
  return ref_models_remap2(wv, n, ids, xyza, layer, 0);
}

// Model::TestCollision() and Model::AppendTouchingModels() for n models, through the interposers above, so that
// every call lands in the trace with its answer (tests that pin the oracle's restatement of the two on worlds
// where plenty of models overlap). Returns how many of the collision tests found something.
int ref_probe_models(void *wv, uint32_t n, const uint32_t *ids)
{
  (void)wv;
  int hits = 0;
  for (uint32_t i = 0; i < n; i++) {
    Model *m = Model::LookupId(ids[i]);
    if (!m)
      continue;
    if (m->TestCollision())
      hits++;
    std::set<Model *> touchers;
    m->AppendTouchingModels(touchers);
  }
  return hits;
}

// ModelPosition::SetSpeed (model_position.cc:595): what a controller does to drive a robot.
int ref_model_set_speed(void *wv, const char *name, double x, double y, double a)
{
  World *w = (World *)wv;
  ModelPosition *m = dynamic_cast<ModelPosition *>(w->GetModel(std::string(name)));
  if (!m)
    return -1;
  m->SetSpeed(x, y, a);
  return 0;
}

// Keep every fiducial finder on the main thread's queue from now on (call BEFORE loading a world).
// On worker threads ModelFiducial::Update reads poses that ModelPosition::Move is changing
// concurrently (world.cc:647 vs model_fiducial.cc:124-127), so recordings would not be reproducible
// (SURVEY.md 3.1). Done by answering World::GetEventQueue (world.cc:678-686, asked by
// Model::Startup, model.cc:577) with queue 0 for them.
void ref_fiducials_on_main_thread(int on) { g_fiducials_on_main = on != 0; }

// Every block of every model as the reference holds it right now: what stg_rt_block_define / stg_rt_models_pose_update
// take (the block's polygon in model coordinates after BlockGroup::CalcSize, Block::local_z, the model's frame
// GetGlobalPose() + geom.pose) and what Model::LocalToPixels (model.cc:474-491) makes of them - the reference's own answer
// for the device's pose path to be held to. Models in ascending id order, blocks in BlockGroup order.
//   hdr:  int64[4] per block {model id, n_pts, first point (index into pts / pix), 0}
//   num:  double[6] per block {gpose x, y, z, a, local_z min, max}
//   pts:  double[2] per point (Block::pts);  pix: int32[2] per point (LocalToPixels)
// Returns the number of blocks (may exceed cap_blocks; n_points gets the total point count). Any out array may be NULL.
int64_t ref_block_geometry(void *wv, int64_t *hdr, double *num, int64_t cap_blocks, double *pts, int32_t *pix, int64_t cap_points,
                           int64_t *n_points)
{
  World *w = (World *)wv;
  std::vector<Model *> models(w->models.begin(), w->models.end());
  std::sort(models.begin(), models.end(), [](const Model *a, const Model *b) { return a->id < b->id; });
  int64_t nb = 0, np = 0;
  for (size_t i = 0; i < models.size(); i++) {
    Model *m = models[i];
    const Pose gp = m->GetGlobalPose() + m->geom.pose;
    for (size_t k = 0; k < m->blockgroup.blocks.size(); k++) {
      const Block &b = m->blockgroup.blocks[k];
      const std::vector<point_int_t> px = m->LocalToPixels(b.pts);
      if (nb < cap_blocks && hdr && num) {
        hdr[4 * nb] = m->id;
        hdr[4 * nb + 1] = (int64_t)b.pts.size();
        hdr[4 * nb + 2] = np;
        hdr[4 * nb + 3] = 0;
        const double d[6] = { gp.x, gp.y, gp.z, gp.a, b.local_z.min, b.local_z.max };
        memcpy(num + 6 * nb, d, sizeof(d));
      }
      for (size_t j = 0; j < b.pts.size(); j++, np++)
        if (np < cap_points && pts && pix) {
          pts[2 * np] = b.pts[j].x;
          pts[2 * np + 1] = b.pts[j].y;
          pix[2 * np] = px[j].x;
          pix[2 * np + 1] = px[j].y;
        }
      nb++;
    }
  }
  if (n_points)
    *n_points = np;
  return nb;
}

// Look a model up by name (World::GetModel, world.cc) -> Model::id, or -1.
int32_t ref_model_id(void *wv, const char *name)
{
  World *w = (World *)wv;
  Model *m = w->GetModel(std::string(name));
  return m ? (int32_t)m->id : -1;
}

// Global pose of a model (Model::GetGlobalPose, model.cc:1205-1218).
int ref_model_pose(void *wv, uint32_t id, double *xyza)
{
  World *w = (World *)wv;
  for (std::set<Model *>::iterator it = w->models.begin(); it != w->models.end(); ++it)
    if ((*it)->id == id) {
      Pose p = (*it)->GetGlobalPose();
      xyza[0] = p.x;
      xyza[1] = p.y;
      xyza[2] = p.z;
      xyza[3] = p.a;
      return 0;
    }
  return -1;
}

// Copy a ranger sensor's latest scan (ModelRanger::Sensor, stage.hh:2643-2645). Returns
// sample_count, or -1. Any of the output pointers may be NULL.
int ref_ranger_scan(void *wv, uint32_t id, int sensor, double *ranges, double *intensities,
                    double *bearings, int cap)
{
  World *w = (World *)wv;
  for (std::set<Model *>::iterator it = w->models.begin(); it != w->models.end(); ++it)
    if ((*it)->id == id) {
      ModelRanger *r = dynamic_cast<ModelRanger *>(*it);
      if (!r || sensor >= (int)r->GetSensors().size())
        return -1;
      const ModelRanger::Sensor &s = r->GetSensors()[sensor];
      const int n = (int)s.ranges.size();
      for (int i = 0; i < n && i < cap; i++) {
        if (ranges)
          ranges[i] = s.ranges[i];
        if (intensities)
          intensities[i] = s.intensities[i];
        if (bearings)
          bearings[i] = s.bearings[i];
      }
      return n;
    }
  return -1;
}

} // extern "C"
