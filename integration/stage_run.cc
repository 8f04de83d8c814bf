// stage_run - a headless front end like libstage/main.cc's `stage -g`, for comparing and timing runs: loads a world
// file with its controllers, runs a fixed number of World::Update()s and, when asked, writes after every tick what the
// controllers could read - every ranger sensor's ranges / intensities / bearings, every fiducial finder's records,
// every model's global pose - to a binary dump. tests/test_gpu_libstage.py runs it twice, plain and with
// LD_PRELOAD=libstage_gpu.so STAGE_GPU=1, and compares the dumps byte for byte.
//
//   stage_run <world file> <ticks> [dump file]
//   STAGE_FIDUCIALS_ON_MAIN=1   keep fiducial finders on the main thread's queue: on worker threads
//                               ModelFiducial::Update reads poses that ModelPosition::Move is changing at the same time
//                               (world.cc:647 vs model_fiducial.cc:124-127), so not even two runs of the unmodified
//                               reference agree
// Uses only Stage's public API, plus one interposed member (World::GetEventQueue) for the switch above.
// TEST / MEASUREMENT TOOL: built by integration/build.sh next to the reference build, never part of libstg_rt.so.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <dlfcn.h>

#include "stage.hh"

using namespace Stg;

// World::GetEventQueue (world.cc:678-686), asked by Model::Startup (model.cc:577) for thread_safe models
unsigned int World::GetEventQueue(Model *mod) const
{
  typedef unsigned int (*fn_t)(const World *, Model *);
  static fn_t fn = (fn_t)dlsym(RTLD_NEXT, "_ZNK3Stg5World13GetEventQueueEPNS_5ModelE");
  static const bool on_main = getenv("STAGE_FIDUCIALS_ON_MAIN") && atoi(getenv("STAGE_FIDUCIALS_ON_MAIN"));
  if (on_main && dynamic_cast<ModelFiducial *>(mod))
    return 0;
  return fn(this, mod);
}

// debugging aid (STAGE_RUN_TRACE_RNG=1): every random() a controller draws, with the tick it is drawn in
static long g_tick = -1;
extern "C" long random(void)
{
  typedef long (*fn_t)(void);
  static fn_t fn = (fn_t)dlsym(RTLD_NEXT, "random");
  static const bool trace = getenv("STAGE_RUN_TRACE_RNG") != NULL;
  const long v = fn();
  if (trace && g_tick >= 0)
    fprintf(stderr, "random() in tick %ld -> %ld\n", g_tick, v);
  return v;
}

// debugging aid (STAGE_RUN_TRACE_VEL=1): every turn-speed command a controller gives
void ModelPosition::SetTurnSpeed(double a)
{
  typedef void (*fn_t)(ModelPosition *, double);
  static fn_t fn = (fn_t)dlsym(RTLD_NEXT, "_ZN3Stg13ModelPosition12SetTurnSpeedEd");
  static const bool trace = getenv("STAGE_RUN_TRACE_VEL") != NULL;
  if (trace)
    fprintf(stderr, "turn tick %ld model %u: %.17g\n", g_tick, GetId(), a);
  fn(this, a);
}

static void put(FILE *f, const void *p, size_t n) { if (f) fwrite(p, 1, n, f); }
static void put_u32(FILE *f, uint32_t v) { put(f, &v, 4); }
static void put_pose(FILE *f, const Pose &p) { const double d[4] = { p.x, p.y, p.z, p.a }; put(f, d, sizeof(d)); }

static bool by_id(const Model *a, const Model *b) { return a->GetId() < b->GetId(); }

int main(int argc, char **argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: stage_run <world file> <ticks> [dump file]\n");
    return 2;
  }
  const long ticks = atol(argv[2]);
  FILE *dump = argc > 3 ? fopen(argv[3], "wb") : NULL;
  Stg::Init(&argc, &argv);
  // Stg::Init seeds drand48 from the clock (stage.cc:27); controllers draw from it (pioneer_flocking.cc:96) and from
  // random() (wander.cc:139): fix both so that two runs of the same world can be compared at all
  srand48(12345);
  srandom(12345);
  World *world = new World(argv[1]); // never destroyed: ~World crashes in the reference itself
  world->Load(argv[1]);
  std::vector<Model *> models;
  {
    const std::set<Model *> all = world->GetAllModels();
    models.assign(all.begin(), all.end());
    std::sort(models.begin(), models.end(), by_id);
  }
  // STAGE_RUN_SUBSCRIBE_ALL=1: every ranger, fiducial finder and blobfinder is subscribed to, as a client or the GUI's data
  // views would: worlds without controllers (lsp_test.world) then update their sensors every tick
  if (getenv("STAGE_RUN_SUBSCRIBE_ALL") && atoi(getenv("STAGE_RUN_SUBSCRIBE_ALL")))
    for (size_t i = 0; i < models.size(); i++)
      if (dynamic_cast<ModelRanger *>(models[i]) || dynamic_cast<ModelFiducial *>(models[i]) || dynamic_cast<ModelBlobfinder *>(models[i]))
        models[i]->Subscribe();
  unsigned long long rays = 0, fid_records = 0, fid_updates = 0, blob_records = 0;
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double in_update = 0;
  for (long t = 0; t < ticks; t++) {
    const std::chrono::steady_clock::time_point a = std::chrono::steady_clock::now();
    g_tick = t;
    const bool done = world->Update();
    g_tick = -1;
    in_update += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    if (done)
      break;
    if (getenv("STAGE_RUN_TRACE_VEL"))
      for (size_t i = 0; i < models.size(); i++)
        if (ModelPosition *p = dynamic_cast<ModelPosition *>(models[i])) {
          const Velocity v = p->GetVelocity();
          fprintf(stderr, "vel tick %ld model %u: %.17g %.17g stall %d\n", t, p->GetId(), v.x, v.a, (int)p->Stalled());
        }
    put_u32(dump, 0x5449434bu); // "TICK"
    put_u32(dump, (uint32_t)t);
    if (getenv("STAGE_RUN_TRACE_RNG")) { // where libc's random() stands after this tick (debugging aid: stderr, not the dump)
      static char other[128];
      static bool ready = false;
      if (!ready) {
        char *cur = initstate(1u, other, sizeof(other));
        setstate(cur);
        ready = true;
      }
      char save[128];
      char *cur = setstate(other);
      memcpy(save, cur, 128);
      setstate(cur);
      const long v = random();
      cur = setstate(other);
      memcpy(cur, save, 128);
      setstate(cur);
      fprintf(stderr, "rngtrace %ld %ld\n", t, v);
    }
    for (size_t i = 0; i < models.size(); i++) {
      Model *m = models[i];
      ModelRanger *r = dynamic_cast<ModelRanger *>(m);
      ModelFiducial *fm = r ? NULL : dynamic_cast<ModelFiducial *>(m);
      ModelBlobfinder *bm = (r || fm) ? NULL : dynamic_cast<ModelBlobfinder *>(m);
      put_u32(dump, m->GetId());
      put_u32(dump, r ? 1u : fm ? 2u : bm ? 3u : 0u); // record kind: what follows the pose
      put_pose(dump, m->GetGlobalPose());
      if (r) {
        const std::vector<ModelRanger::Sensor> &s = r->GetSensors();
        put_u32(dump, (uint32_t)s.size());
        for (size_t k = 0; k < s.size(); k++) {
          put_u32(dump, (uint32_t)s[k].ranges.size());
          put(dump, s[k].ranges.empty() ? NULL : &s[k].ranges[0], s[k].ranges.size() * 8);
          put(dump, s[k].intensities.empty() ? NULL : &s[k].intensities[0], s[k].intensities.size() * 8);
          put(dump, s[k].bearings.empty() ? NULL : &s[k].bearings[0], s[k].bearings.size() * 8);
          rays += s[k].ranges.size();
        }
      } else if (fm) {
        const std::vector<ModelFiducial::Fiducial> &fs = fm->GetFiducials();
        put_u32(dump, (uint32_t)fs.size());
        fid_updates++;
        for (size_t k = 0; k < fs.size(); k++) {
          const double d[2] = { fs[k].range, fs[k].bearing };
          put(dump, d, sizeof(d));
          put_pose(dump, fs[k].geom);
          put_pose(dump, fs[k].pose);
          put_u32(dump, fs[k].mod ? fs[k].mod->GetId() : 0xffffffffu);
          put_u32(dump, (uint32_t)fs[k].id);
          fid_records++;
        }
      } else if (bm) {
        const std::vector<ModelBlobfinder::Blob> &bs = bm->GetBlobs();
        put_u32(dump, (uint32_t)bs.size());
        for (size_t k = 0; k < bs.size(); k++) {
          const double d[5] = { bs[k].color.r, bs[k].color.g, bs[k].color.b, bs[k].color.a, bs[k].range };
          put(dump, d, sizeof(d));
          put_u32(dump, bs[k].left);
          put_u32(dump, bs[k].top);
          put_u32(dump, bs[k].right);
          put_u32(dump, bs[k].bottom);
          blob_records++;
        }
      }
    }
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (dump)
    fclose(dump);
  unsigned long long gpu[13] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  typedef void (*cnt_t)(unsigned long long *);
  if (cnt_t cnt = (cnt_t)dlsym(RTLD_DEFAULT, "stage_gpu_counters"))
    cnt(gpu);
  // one line, machine-readable (the dumps carry the data; rays = ranger samples held by the sensors, summed over ticks)
  printf("\nstage_run: {\"world\": \"%s\", \"ticks\": %llu, \"seconds_in_update\": %.6f, \"seconds_wall\": %.6f, \"ticks_per_s\": %.3f, "
         "\"ranger_samples\": %llu, \"fiducial_records\": %llu, \"blob_records\": %llu, \"gpu\": {\"fans\": %llu, \"fiducial_updates\": %llu, \"map_calls\": %llu, "
         "\"ticks\": %llu, \"state\": \"%s\", \"verified_beams\": %llu, \"mismatched_beams\": %llu, \"verified_fiducials\": %llu, "
         "\"mismatched_fiducial_updates\": %llu, \"fiducial_max_abs_err\": %.3e, \"blob_updates\": %llu, \"verified_blobs\": %llu, "
         "\"mismatched_blob_updates\": %llu}}\n",
         argv[1], (unsigned long long)world->GetUpdateCount(), in_update, wall, world->GetUpdateCount() / in_update, rays, fid_records, blob_records, gpu[0], gpu[1],
         gpu[2], gpu[3], gpu[4] == 4 ? "emulated on the cpu" : gpu[4] == 3 ? "gpu, verified against libstage's own loops in place" : gpu[4] == 2 ? "gpu" : gpu[4] == 1 ? "fell back to cpu" : "cpu",
         gpu[5], gpu[6], gpu[7], gpu[8], gpu[9] * 1e-18, gpu[10], gpu[11], gpu[12]);
  fflush(stdout);
  _exit(0);
}
