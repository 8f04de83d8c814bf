// Host runtime behind the C ABI of include/stg_rt.h: context, host shadow of the mapped blocks,
// delta coalescing and blob building, fan batching, pinned staging, stream plumbing.
// The compute is in rt_kernels.cu; there is no CPU implementation of any of it here.
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/stg_rt.h"
#include "rt_device.h"

namespace stg {
void launch_rehash(const unsigned long long *src, unsigned long long *dst, uint32_t slot_mask, void *stream);
void launch_grid_hash(const GridView &g, unsigned long long *out, void *stream);
void launch_host_store(void *dst, size_t n_words, unsigned long long seed, void *stream);
void launch_grow(const GridView &from, const GridView &to, void *stream);
}

using namespace stg;

static_assert(sizeof(stg_rt_fan) == sizeof(FanRec), "stg_rt_fan and FanRec must share a layout");

namespace {

thread_local std::string g_create_error;

// What the host remembers of a block. The scalars come first and fit two cache lines: a tick that moves 100,000 bodies by
// pose walks these records four times (queue, sort check, size, write) and, laid out this way, never touches the vectors
// behind them (vec_used says whether any of a layer's four may hold something).
struct alignas(64) BlockShadow {
  bool known = false;
  bool z_dirty = false;
  // what the caller last said (host_*) and what the device grid currently holds (dev_*)
  bool host_mapped[2] = { false, false }, dev_mapped[2] = { false, false }, dirty[2] = { false, false };
  bool keep_seq[2] = { false, false }; // re-map of a block the device already holds: its place in the cell lists stays
  // pose path (stg_rt_models_pose_update): the outline is derived on the device from the block's registered polygon
  bool pose_pending[2] = { false, false }; // the queued Map of this layer is a pose, not an outline
  bool dev_owned[2] = { false, false };    // the device rendered the outline it holds itself (GeomView::pix); dev_xy is empty
  bool vec_used[2] = { false, false };     // host_xy / host_poly / dev_xy / dev_poly of the layer may be non-empty
  uint32_t model = 0;
  uint32_t map_seq[2] = { 0, 0 };
  uint32_t host_steps[2] = { 0, 0 }, dev_steps[2] = { 0, 0 }; // cell visits of host_xy / dev_xy (all renderings)
  double zmin = 0, zmax = 0;
  double pose[2][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } };
  std::vector<int32_t> host_xy[2], dev_xy[2];
  // Block::Map on a block that is already mapped renders it once more (block.cc:170-181 does not check; the
  // reference's world loader does it to bumper models): the cells get duplicate entries, Region::count counts them,
  // UnMap removes them all. Then the outline is several polygons back to back and these hold the points of each
  // (empty = one polygon, the normal case).
  std::vector<uint32_t> host_poly[2], dev_poly[2];
};

struct GeomShadow { // stg_rt_block_define: host copy of BlockGeom plus what the host needs to reason without the pixels
  bool defined = false;
  uint32_t first_pt = 0, n_pts = 0, model = 0;
  double zmin = 0, zmax = 0;
  double radius = 0;      // farthest vertex from the model origin: bounds the outline at any pose
  uint32_t est_steps = 0; // upper estimate of the outline's cell visits at any pose
};

struct Submit {
  uint64_t first_beam, n_beams;
  stg_rt_fan_out out;
};

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
};

// A palette index no model has: what the streamed expansion leaves behind in the staging bytes it has read, so that a byte
// still holding it after its stretch's flag has been seen can only be one the march's store has not reached the host with
// yet - the expansion waits for it. The flag protocol (one system-wide fence per stretch, FanBatchView) orders the stores by
// the memory model's fence cumulativity; this makes the delivery right even where that did not hold.
constexpr uint32_t kPaletteUnwritten = 255;

struct FanArrays { // device arrays of one batch / resident set
  DevBuf fans, first, headings_in, trig, heading, trigw; // trigw: per-beam (cos, sin) computed on the device for the refill march
  DevBuf ranges, intens, bearings, hit_model, hit_cell, steps;
};

} // namespace

// A few host threads for the per-beam passes stg_rt_sync makes over large result arrays (expanding intensity palette
// indices into doubles: 8 MB of stores per million beams, more than one core writes in the time the GPU takes to trace
// them). Started on first use; STG_RT_HOST_THREADS sets the count (default: half the cores, at most 8; 1 = inline).
class HostPool {
public:
  ~HostPool()
  {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      stop_flag_.store(true);
    }
    cv_.notify_all();
    for (std::thread &t : threads_)
      t.join();
  }
  void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)> &fn)
  {
    start();
    const size_t parts = std::min<size_t>(threads_.size() + 1, std::max<size_t>(1, n / std::max<size_t>(grain, 1)));
    if (parts <= 1) {
      fn(0, n);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_ = n;
      parts_ = parts;
      next_ = 1; // part 0 is the caller's
      pending_ = parts - 1;
      generation_++;
      gen_.store(generation_, std::memory_order_release);
    }
    cv_.notify_all();
    fn(0, n / parts);
    for (int spin = 0; spin < 20000 && pend_peek() != 0; spin++) // the others are about as long as this part was
      _mm_pause();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

  // The same without waiting: the parts are taken by the workers as they come free; finish() takes what is left itself
  // and returns when every part is done. One job at a time.
  void begin(size_t parts, std::function<void(size_t)> part_fn)
  {
    start();
    {
      std::lock_guard<std::mutex> lk(mu_);
      async_fn_ = std::move(part_fn);
      wrapped_ = [this](size_t lo, size_t hi) {
        for (size_t p = lo; p < hi; p++)
          async_fn_(p);
      };
      fn_ = &wrapped_;
      n_ = parts;
      parts_ = parts;
      next_ = 0;
      pending_ = parts;
      generation_++;
      gen_.store(generation_, std::memory_order_release);
      async_open_ = true;
    }
    cv_.notify_all();
  }
  void finish()
  {
    std::unique_lock<std::mutex> lk(mu_);
    if (!async_open_)
      return;
    while (next_ < parts_) {
      const size_t p = next_++;
      lk.unlock();
      async_fn_(p);
      lk.lock();
      if (--pending_ == 0)
        done_.notify_all();
    }
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
    async_open_ = false;
  }
  bool busy()
  {
    std::lock_guard<std::mutex> lk(mu_);
    return async_open_;
  }

private:
  std::function<void(size_t)> async_fn_;
  std::function<void(size_t, size_t)> wrapped_;
  bool async_open_ = false;
  void start()
  {
    if (started_)
      return;
    started_ = true;
    unsigned want = std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    if (const char *e = getenv("STG_RT_HOST_THREADS"))
      want = (unsigned)std::max(1, atoi(e));
    for (unsigned i = 1; i < want; i++)
      threads_.emplace_back([this] { work(); });
  }
  void work()
  {
    uint64_t seen = 0;
    for (;;) {
      // Ticks follow each other within a millisecond: a worker that has just finished polls for the next job for a
      // while before it goes to sleep (waking a sleeping thread costs more than the job it is woken for).
      const std::chrono::steady_clock::time_point until = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
      while (gen_.load(std::memory_order_acquire) == seen && !stop_flag_.load(std::memory_order_relaxed) &&
             std::chrono::steady_clock::now() < until)
        _mm_pause();
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return stop_ || (generation_ != seen && next_ < parts_) || generation_ != seen; });
      if (stop_)
        return;
      while (fn_ && next_ < parts_) {
        const size_t p = next_++;
        const size_t lo = n_ * p / parts_, hi = n_ * (p + 1) / parts_;
        const std::function<void(size_t, size_t)> *fn = fn_;
        lk.unlock();
        (*fn)(lo, hi);
        lk.lock();
        if (--pending_ == 0)
          done_.notify_all();
      }
      seen = generation_;
    }
  }
  size_t pend_peek()
  {
    std::lock_guard<std::mutex> lk(mu_);
    return pending_;
  }
  std::atomic<uint64_t> gen_{ 0 };
  std::atomic<bool> stop_flag_{ false };
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> threads_;
  const std::function<void(size_t, size_t)> *fn_ = nullptr;
  size_t n_ = 0, parts_ = 0, next_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  bool started_ = false, stop_ = false;
};

struct stg_rt_set {
  std::vector<FanRec> fans;
  std::vector<uint32_t> first;
  uint64_t n_beams = 0;
  uint32_t want = 0;
  bool has_trig = false, has_headings = false;
  FanArrays d;
  bool fans_dirty = true;
};

struct stg_rt_ctx {
  std::mutex mu;
  stg_rt_config cfg;
  std::string err;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  GridView g;
  size_t n_regions = 0;
  uint32_t slot_cap = 0;
  unsigned long long *spare_slots = nullptr;
  int32_t cell_x0 = 0, cell_y0 = 0, cell_x1 = 0, cell_y1 = 0; // inclusive cell extent

  std::vector<BlockShadow> blocks;
  std::vector<uint32_t> dirty_units; // block*2 + layer, in order of first touch
  bool units_in_order = true;        // ... which so far is also ascending order of sort_units' key (touch_unit)
  uint32_t last_unit_key = 0;
  std::vector<ModelUpd> model_upd;
  std::vector<bool> model_known;
  std::vector<ModelUpd> models;          // host copy of the model table
  bool block_reassigned = false;         // a known block id was mapped for another model
  std::vector<uint32_t> models_changed;  // models whose root / visibility changed while blocks are mapped
  uint64_t epoch = 0;
  uint32_t local_seq = 0;
  int last_export_rank = -1; // multi-GPU: this context's rank, learnt from stg_rt_delta_export
  bool export_mode = false;  // deltas leave through stg_rt_delta_export (never applied locally)
  int64_t last_export_ins[2] = { 0, 0 }, last_gather_ins[2] = { 0, 0 }, last_export_rem[2] = { 0, 0 }, last_gather_rem[2] = { 0, 0 };
  PinBuf h_hdrs;
  cudaEvent_t hdrs_ready = nullptr, xstream_ev = nullptr;
  bool hdrs_pending = false;
  int hdrs_ranks = 0;
  uint32_t *err_host = nullptr;       // page-locked word the kernels raise kErr* bits in (GridView::err_host)
  int64_t live_entries[2] = { 0, 0 }; // cell visits currently rendered, per layer
  int64_t used_upper[2] = { 0, 0 };   // upper bound of non-empty slots (entries + tombstones)

  std::vector<uint32_t> attr_blocks;
  // registered block geometry (pose path)
  std::vector<GeomShadow> geom;
  std::vector<double> h_pts;                      // pool of local points, 2 doubles each
  std::vector<std::vector<uint32_t> > model_blocks; // per model: its defined blocks in definition (= BlockGroup) order
  DevBuf d_geom, d_pts, d_pix[2], d_pix_new[2];
  DevBuf d_work_counter; // the refill march's chunk counter
  bool geom_dirty = false;
  GeomView gv;
  PinBuf h_blob;
  DevBuf d_blob;
  size_t blob_cap = 0;
  cudaEvent_t blob_uploaded = nullptr; // the pinned blob may be rewritten once this has fired
  cudaEvent_t t_ev[6][2] = {};          // [deltas, headings, march, fiducials, blobs, collisions / touchers][start, stop]
  cudaStream_t copy_stream = nullptr;   // result copies of chunk i overlap the march of chunk i+1
  // k1_summaries runs here after every K1 pass, beside whatever the main stream does next (the march may read either
  // version of a summary word); the next pass waits for it
  cudaStream_t sum_stream = nullptr;
  cudaEvent_t sum_ready = nullptr, sum_done = nullptr;
  bool sum_pending = false;
  cudaEvent_t chunk_ev[8] = {};
  cudaEvent_t copies_done = nullptr;
  bool t_rec[6] = { false, false, false, false, false, false };

  // fans queued by stg_rt_fans_submit and not launched yet
  // the queued fan records live in page-locked memory and are uploaded from where the caller's records were
  // copied to (a batch of sonars is a million records: every extra pass over them costs milliseconds)
  struct PinFans {
    FanRec *p = nullptr;
    size_t n = 0, cap = 0;
    size_t size() const { return n; }
    FanRec *data() { return p; }
    void clear() { n = 0; }
  } q_fans;
  cudaEvent_t fans_uploaded = nullptr;
  // a queued stg_rt_rangers_submit: the fan records are derived on the device from these
  struct RigJob {
    bool active = false;
    uint32_t n_sensors = 0, n_rangers = 0, uniform_samples = 0;
    size_t o_first = 0, o_poses = 0, bytes = 0; // layout of h_rig / d_rig: sensors | sensor_first | poses
  } q_rig;
  PinBuf h_rig;
  DevBuf d_rig;
  std::vector<uint32_t> q_first;
  std::vector<double> q_headings, q_trig;
  std::vector<FanNoiseRec> q_noise; // parallel to q_fans once any noisy fan is queued
  bool q_has_noise = false;
  uint64_t noise_epoch = 0;
  std::vector<Submit> q_subs;
  uint64_t q_beams = 0;
  bool q_has_trig = false;
  // launched, results not delivered yet
  std::vector<Submit> f_subs;
  uint64_t f_beams = 0;
  bool f_direct[3] = { false, false, false }; // ranges / intensities / hit_model written in place
  bool f_copied = false;                       // ... or fetched with the copy engine (STG_RT_RESULTS=copy)
  FanArrays d;
  DevBuf d_in; // fans | first_beam | headings | trig of the batch in flight
  PinBuf h_in, h_ranges, h_intens, h_bearings, h_hit_model, h_hit_cell, h_steps;

  // sensor post-processing batches (fiducial finders, blobfinders): one in flight per kind
  struct PostJob {
    void *user_out = nullptr;
    uint32_t *user_count = nullptr;
    size_t out_bytes = 0, count_bytes = 0;
    bool pending = false;
    uint32_t n_finders = 0, max_per_finder = 0, n_targets = 0; // fiducials: to unpack the packed records
    uint64_t packed_capacity = 0;               // ... or, non-zero, to hand them over packed
  } fid_job, blob_job;
  DevBuf d_fid_grid, d_fid_in, d_fid_out, d_fid_count, d_fid_off, d_blob_in, d_blob_out, d_blob_count, d_mdl_color;
  PinBuf h_fid_in, h_fid_out, h_fid_count, h_blob_in, h_blob_out, h_blob_count;
  FanArrays bd; // the blobfinders' scan-line fans
  // STG_RT_HOST_PROF=1: host microseconds per section of the flush path, printed by stg_rt_destroy
  double prof_us[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  uint64_t prof_flushes = 0;
  bool prof_on = getenv("STG_RT_HOST_PROF") != nullptr;
  struct CollisionJob {
    int32_t *user_hit = nullptr;
    std::vector<int32_t> fallback; // per query: what to report when no cell test hits (-1, or the ground model)
    bool pending = false;
  } col_job;
  DevBuf d_col_in, d_col_hit;
  PinBuf h_col_in, h_col_hit;
  // stg_rt_touching_models
  struct ToucherJob {
    uint32_t *user_out = nullptr, *user_count = nullptr;
    uint32_t n_queries = 0, max_per_query = 0;
    bool pending = false;
  } touch_job;
  DevBuf d_touch_out;
  PinBuf h_touch_out;
  // stg_rt_models_move
  DevBuf d_mv_in, d_mv_out, d_mover_of_model;
  PinBuf h_mv_in, h_mv_out;
  bool hold_epoch = false; // the undo re-maps of a move belong to the same epoch as the moves
  std::vector<double> mdl_color; // rgba per model
  bool mdl_color_dirty = true;
  // the distinct ranger_return values of the world (index 0 = "no hit" = 0.0): with at most 254 of them a beam's
  // intensity crosses PCIe as one byte and stg_rt_sync expands it into the caller's doubles
  std::vector<double> palette = std::vector<double>(1, 0.0);
  bool palette_ok = true;
  size_t idx_clean_bytes = 0; // so many leading bytes of h_intens_idx read kPaletteUnwritten (the streamed expansion puts them back as it goes)
  const void *idx_clean_base = nullptr; // ... of the buffer at this address
  bool f_compact_intens = false; // the batch in flight delivers intensities as palette indices
  PinBuf h_intens_idx;
  // streamed delivery: the march tells the host, stretch by stretch, which results have landed (FanBatchView::stretch_done;
  // d_stretch_count = its per-stretch warp counters), and the pool's threads expand intensities of finished stretches while
  // the kernel is still running
  PinBuf h_done_flags;
  DevBuf d_stretch_count;
  uint32_t done_seq = 0;
  bool f_streaming = false;
  std::atomic<bool> stream_gave_up{ false };
  HostPool pool;

  std::map<char *, size_t> host_allocs; // stg_rt_host_alloc regions
  stg_rt_stats stats;

  stg_rt_ctx() { memset(&stats, 0, sizeof(stats)); memset(&g, 0, sizeof(g)); memset(&gv, 0, sizeof(gv)); }
};

namespace {

int fail(stg_rt_ctx *c, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c)
    c->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define CU(ctx, call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, STG_RT_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int summaries_fence_locked(stg_rt_ctx *c);
int summaries_launch_locked(stg_rt_ctx *c);

int dev_reserve(stg_rt_ctx *c, DevBuf &b, size_t bytes)
{
  if (bytes <= b.cap)
    return 0;
  if (b.p)
    CU(c, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  const size_t want = std::max(bytes + bytes / 4, (size_t)4096);
  CU(c, cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}

int pin_reserve(stg_rt_ctx *c, PinBuf &b, size_t bytes)
{
  if (bytes <= b.cap)
    return 0;
  if (b.p)
    CU(c, cudaFreeHost(b.p));
  b.p = nullptr;
  b.cap = 0;
  const size_t want = std::max(bytes + bytes / 4, (size_t)4096);
  CU(c, cudaMallocHost(&b.p, want));
  b.cap = want;
  return 0;
}

void free_fan_arrays(FanArrays &a)
{
  DevBuf *all[] = { &a.fans, &a.first, &a.headings_in, &a.trig, &a.heading, &a.trigw, &a.ranges,
                    &a.intens, &a.bearings, &a.hit_model, &a.hit_cell, &a.steps };
  for (DevBuf *b : all) {
    if (b->p)
      cudaFree(b->p);
    b->p = nullptr;
    b->cap = 0;
  }
}

uint32_t next_pow2(uint64_t v)
{
  uint64_t p = 1;
  while (p < v)
    p <<= 1;
  return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

bool is_host_alloc(stg_rt_ctx *c, const void *p, size_t bytes)
{
  if (!p || c->host_allocs.empty())
    return false;
  auto it = c->host_allocs.upper_bound((char *)p);
  if (it == c->host_allocs.begin())
    return false;
  --it;
  return (char *)p >= it->first && (char *)p + bytes <= it->first + it->second;
}

// ---- deltas -----------------------------------------------------------------------------------

uint32_t root_flags_of(const stg_rt_ctx *c, uint32_t model)
{
  const ModelUpd &m = c->models[model];
  return (m.root & kRecRootMask) | (m.ranger_return < 0 ? kRecNoRanger : 0u) |
         ((m.flags & STG_RT_VIS_GRIPPER_RETURN) ? kRecGripper : 0u) | ((m.flags & STG_RT_VIS_OBSTACLE_RETURN) ? kRecObstacle : 0u);
}

// Units sorted so that the blob (and, when it has to be split, successive blobs) insert in Map-call
// order: pure removals first, then by the sequence number of the final Map.
void sort_units(stg_rt_ctx *c)
{
  auto key = [c](uint32_t u) {
    const BlockShadow &b = c->blocks[u >> 1];
    return b.host_mapped[u & 1] ? b.map_seq[u & 1] : 0u;
  };
  auto less = [&key](uint32_t a, uint32_t b) { return key(a) < key(b); };
  if (c->units_in_order) // queued in ascending key order (touch_unit kept track): nothing to do, nothing to read
    return;
  if (!std::is_sorted(c->dirty_units.begin(), c->dirty_units.end(), less)) // usually already in Map-call order
    std::stable_sort(c->dirty_units.begin(), c->dirty_units.end(), less);
}

// Keep probe chains short: tombstones never turn back into empty slots on their own, so when the
// upper bound of non-empty slots (entries + tombstones) would pass 3/4 of the table, re-insert the
// live entries into the spare table and swap. Runs between the removal and the insertion phase of a
// blob, when the table holds exactly live_after - add[L] entries (live_entries is already updated).
int ensure_table_room(stg_rt_ctx *c, const int64_t add[2])
{
  for (int L = 0; L < 2; L++) {
    if (c->used_upper[L] + add[L] > (int64_t)c->slot_cap * 3 / 4) {
      CU(c, cudaMemsetAsync(c->spare_slots, 0, (size_t)c->slot_cap * 8, c->stream));
      launch_rehash(c->g.slots[L], c->spare_slots, c->g.slot_mask, c->stream);
      c->stats.launches_other++;
      std::swap(c->g.slots[L], c->spare_slots);
      c->used_upper[L] = c->live_entries[L] - add[L];
    }
    c->used_upper[L] += add[L];
  }
  return 0;
}

// Emit the queued deltas as one or more blobs. apply == true: upload and apply each blob locally.
// apply == false (multi-GPU export): exactly one blob, uploaded to d_blob, not applied.
// Write the edge records of one outline straight into the blob. Zero-length edges are kept (they
// own no step, so the kernels' edge search never lands on them).
static inline uint32_t emit_polygon(const int32_t *p, size_t n, uint32_t block, uint32_t layer, uint32_t &step_cursor, EdgeRec *&dst)
{
  uint32_t steps = 0;
  for (size_t i = 0; i < n; i++) {
    const size_t j = i + 1 == n ? 0 : i + 1;
    EdgeRec &e = *dst++;
    e.x0 = p[2 * i];
    e.y0 = p[2 * i + 1];
    e.x1 = p[2 * j];
    e.y1 = p[2 * j + 1];
    const int64_t adx = e.x1 > e.x0 ? (int64_t)e.x1 - e.x0 : (int64_t)e.x0 - e.x1;
    const int64_t ady = e.y1 > e.y0 ? (int64_t)e.y1 - e.y0 : (int64_t)e.y0 - e.y1;
    e.n_steps = (uint32_t)(adx + ady);
    e.block = block;
    e.layer = layer;
    e.first_step = step_cursor;
    step_cursor += e.n_steps;
    steps += e.n_steps;
  }
  return steps;
}

static inline uint32_t emit_edges(const std::vector<int32_t> &xy, uint32_t block, uint32_t layer, uint32_t &step_cursor,
                                  EdgeRec *&dst, const std::vector<uint32_t> *polys = nullptr)
{
  if (!polys || polys->empty())
    return emit_polygon(xy.data(), xy.size() / 2, block, layer, step_cursor, dst);
  uint32_t steps = 0;
  size_t at = 0;
  for (uint32_t n : *polys) { // a block rendered more than once: every rendering is a closed outline of its own
    steps += emit_polygon(xy.data() + 2 * at, n, block, layer, step_cursor, dst);
    at += n;
  }
  return steps;
}

static inline double prof_now(const stg_rt_ctx *c)
{
  if (!c->prof_on)
    return 0;
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Bring the device copy of the registered block geometry up to date (after stg_rt_block_define calls) and make sure
// the per-layer pixel pools cover it. The pools grow by reallocation; what they hold is copied over.
int upload_geometry_locked(stg_rt_ctx *c)
{
  if (!c->geom_dirty)
    return 0;
  const size_t n_blocks = c->geom.size(), n_vals = c->h_pts.size(); // values = 2 per point
  std::vector<BlockGeom> recs(n_blocks);
  for (size_t b = 0; b < n_blocks; b++) {
    const GeomShadow &g = c->geom[b];
    memset(&recs[b], 0, sizeof(BlockGeom));
    recs[b].first_pt = g.first_pt;
    recs[b].n_pts = g.defined ? g.n_pts : 0;
    recs[b].model = g.model;
    recs[b].est_steps = g.est_steps;
    recs[b].zmin = g.zmin;
    recs[b].zmax = g.zmax;
  }
  int rc;
  if ((rc = dev_reserve(c, c->d_geom, std::max<size_t>(n_blocks * sizeof(BlockGeom), 32))) ||
      (rc = dev_reserve(c, c->d_pts, std::max<size_t>(n_vals * 8, 16))))
    return rc;
  for (int L = 0; L < 2; L++)
    for (DevBuf *pool : { &c->d_pix[L], &c->d_pix_new[L] }) {
      if (n_vals * 4 <= pool->cap)
        continue;
      DevBuf grown;
      if ((rc = dev_reserve(c, grown, std::max<size_t>(n_vals * 4, 16))))
        return rc;
      CU(c, cudaMemsetAsync(grown.p, 0, grown.cap, c->stream));
      if (pool->p) {
        CU(c, cudaMemcpyAsync(grown.p, pool->p, pool->cap, cudaMemcpyDeviceToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaFree(pool->p));
      }
      *pool = grown;
    }
  CU(c, cudaMemcpyAsync(c->d_geom.p, recs.data(), n_blocks * sizeof(BlockGeom), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_pts.p, c->h_pts.data(), n_vals * 8, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream)); // pageable sources
  c->stats.h2d_bytes += n_blocks * sizeof(BlockGeom) + n_vals * 8;
  c->gv.geom = (const BlockGeom *)c->d_geom.p;
  c->gv.pts = (const double *)c->d_pts.p;
  for (int L = 0; L < 2; L++) {
    c->gv.pix[L] = (int32_t *)c->d_pix[L].p;
    c->gv.pix_new[L] = (int32_t *)c->d_pix_new[L].p;
  }
  c->gv.n_blocks = (uint32_t)n_blocks;
  c->geom_dirty = false;
  return 0;
}

int build_and_send_deltas(stg_rt_ctx *c, bool apply, int rank, size_t *out_bytes)
{
  {
    const int rcg = upload_geometry_locked(c);
    if (rcg)
      return rcg;
  }
  double t_prof = prof_now(c);
  auto lap = [&](int k) {
    if (c->prof_on) {
      const double t = prof_now(c);
      c->prof_us[k] += t - t_prof;
      t_prof = t;
    }
  };
  sort_units(c);
  lap(1);
  size_t unit = 0;
  bool first_round = true;
  // model or attribute updates may move a rendered block to another tree: the region owners are then rebuilt
  // on the device (which decides whether that really happened, rt_kernels.cu)
  const bool owners_suspect = !c->model_upd.empty() || !c->models_changed.empty() || c->block_reassigned;
  c->block_reassigned = false;
  if (out_bytes)
    *out_bytes = 0;
  while (first_round || unit < c->dirty_units.size()) {
    // ---- pass 1: how much of the queue fits this blob, and where its sections start ----
    std::vector<uint32_t> &attr_blocks = c->attr_blocks; // blocks whose model's attributes changed (no new Map)
    attr_blocks.clear();
    if (first_round && !c->models_changed.empty()) {
      for (uint32_t bi = 0; bi < c->blocks.size(); bi++) {
        const BlockShadow &b = c->blocks[bi];
        if (b.known && std::find(c->models_changed.begin(), c->models_changed.end(), b.model) != c->models_changed.end())
          attr_blocks.push_back(bi);
      }
      c->models_changed.clear();
    }
    const size_t n_model = first_round ? c->model_upd.size() : 0;
    size_t n_rem = 0, n_ins = 0, n_bupd = attr_blocks.size(), n_pose = 0, end_unit = unit;
    size_t bytes = sizeof(DeltaHeader) + n_bupd * sizeof(BlockUpd) + n_model * sizeof(ModelUpd);
    if (bytes > c->blob_cap)
      return fail(c, STG_RT_ENOMEM, "model / block attribute updates exceed the delta buffer");
    int64_t d_live[2] = { 0, 0 }; // net change of rendered cell visits per layer, known before anything is touched
    // what one unit adds to the blob and to the tables
    struct UnitTally {
      size_t bytes = 0, n_rem = 0, n_ins = 0, n_pose = 0, n_bupd = 0, max_edges = 0;
      int64_t d_live[2] = { 0, 0 };
    };
    auto tally_unit = [c](uint32_t u, UnitTally &t) {
      const BlockShadow &b = c->blocks[u >> 1];
      const int L = u & 1;
      // the outline the device holds goes out through edge records when the host rendered it, through a pose record
      // (the device knows it: GeomView::pix) when the device did; the new one comes in either as edge records or as a pose
      const bool old_by_dev = b.dev_mapped[L] && b.dev_owned[L], new_by_pose = b.host_mapped[L] && b.pose_pending[L];
      const bool new_by_edges = b.host_mapped[L] && !b.pose_pending[L];
      const size_t re = (b.dev_mapped[L] && !b.dev_owned[L]) ? b.dev_xy[L].size() / 2 : 0, ie = new_by_edges ? b.host_xy[L].size() / 2 : 0;
      const size_t pr = (old_by_dev || new_by_pose) ? 1 : 0;
      t.bytes += (re + ie) * sizeof(EdgeRec) + (new_by_edges ? sizeof(BlockUpd) : 0) + pr * sizeof(PoseUpd);
      t.n_rem += re;
      t.n_ins += ie;
      t.n_pose += pr;
      t.n_bupd += new_by_edges ? 1 : 0;
      t.max_edges = re + ie;
      t.d_live[L] += (b.host_mapped[L] ? (int64_t)b.host_steps[L] : 0) - (b.dev_mapped[L] ? (int64_t)b.dev_steps[L] : 0);
    };
    // A tick that moves a swarm queues 100,000 units: the host threads tally them side by side, in kUnitChunks stretches
    // (pass 2 writes the same stretches side by side when there is nothing but pose records in them). If the whole
    // queue fits this blob - the usual case - that is the sizing; otherwise unit by unit up to the one that does not fit.
    constexpr size_t kUnitChunks = 32;
    UnitTally chunk_tally[kUnitChunks];
    bool chunked = false;
    const size_t n_queued = c->dirty_units.size() - unit;
    if (n_queued >= 8192 && !c->pool.busy()) {
      const uint32_t *units = c->dirty_units.data() + unit;
      c->pool.parallel_for(kUnitChunks, 1, [&](size_t k0, size_t k1) {
        for (size_t k = k0; k < k1; k++)
          for (size_t i = k * n_queued / kUnitChunks, e = (k + 1) * n_queued / kUnitChunks; i < e; i++)
            tally_unit(units[i], chunk_tally[k]);
      });
      UnitTally all;
      for (const UnitTally &t : chunk_tally) {
        all.bytes += t.bytes; all.n_rem += t.n_rem; all.n_ins += t.n_ins; all.n_pose += t.n_pose; all.n_bupd += t.n_bupd;
        all.d_live[0] += t.d_live[0]; all.d_live[1] += t.d_live[1];
      }
      if (bytes + all.bytes <= c->blob_cap) {
        chunked = true;
        bytes += all.bytes; n_rem += all.n_rem; n_ins += all.n_ins; n_pose += all.n_pose; n_bupd += all.n_bupd;
        d_live[0] += all.d_live[0]; d_live[1] += all.d_live[1];
        end_unit = c->dirty_units.size();
      }
    }
    while (!chunked && end_unit < c->dirty_units.size()) {
      UnitTally t;
      tally_unit(c->dirty_units[end_unit], t);
      if (bytes + t.bytes > c->blob_cap) {
        if (end_unit == unit)
          return fail(c, STG_RT_ENOMEM, "one block's outline (%zu edges) exceeds the delta buffer", t.max_edges);
        break;
      }
      bytes += t.bytes;
      n_rem += t.n_rem;
      n_ins += t.n_ins;
      n_pose += t.n_pose;
      n_bupd += t.n_bupd;
      d_live[0] += t.d_live[0];
      d_live[1] += t.d_live[1];
      end_unit++;
    }
    if (!apply && end_unit < c->dirty_units.size())
      return fail(c, STG_RT_ENOMEM, "deltas of one tick exceed the all-gather slot (%zu bytes)", c->blob_cap);
    // the cell-entry tables must keep half their slots free; checked here, while the shadow still says what the
    // device holds, so that a caller who gets this error can unmap something and carry on
    for (int L = 0; L < 2; L++)
      if (c->live_entries[L] + d_live[L] > (int64_t)c->slot_cap / 2)
        return fail(c, STG_RT_ENOMEM, "cell-entry table of layer %d full (%lld entries, capacity %u/2); raise max_cell_entries",
                    L, (long long)(c->live_entries[L] + d_live[L]), c->slot_cap);

    lap(2);
    // ---- pass 2: write the records straight into the pinned blob ----
    CU(c, cudaEventSynchronize(c->blob_uploaded)); // the previous upload has left the pinned buffer
    unsigned char *base = (unsigned char *)c->h_blob.p;
    DeltaHeader *hdr = (DeltaHeader *)base;
    EdgeRec *rem = (EdgeRec *)(base + sizeof(DeltaHeader)), *ins = rem + n_rem;
    BlockUpd *bupd = (BlockUpd *)(ins + n_ins);
    ModelUpd *mupd = (ModelUpd *)(bupd + n_bupd);
    PoseUpd *pupd = (PoseUpd *)(mupd + n_model);
    int64_t est_rem[2] = { 0, 0 }, est_ins[2] = { 0, 0 };
    for (uint32_t bi : attr_blocks) {
      const BlockShadow &b = c->blocks[bi];
      BlockUpd &up = *bupd++;
      up.block = bi;
      up.model = b.model;
      up.zmin = b.zmin;
      up.zmax = b.zmax;
      up.seq_local = 0;
      up.layer = kUpdKeepSeq;
      up.root_flags = root_flags_of(c, b.model);
      up.pad = 0;
    }
    uint32_t rem_cursor = 0, ins_cursor = 0;
    int64_t ins_round[2] = { 0, 0 }, rem_round[2] = { 0, 0 };
    // one unit: its records, and the shadow brought to "the device holds what the caller last said"
    auto write_unit = [c, &rem, &ins](uint32_t u, PoseUpd *&pupd, BlockUpd *&bupd, uint32_t &rem_cursor, uint32_t &ins_cursor, int64_t *est_rem,
                                   int64_t *est_ins, int64_t *rem_round, int64_t *ins_round) {
      BlockShadow &b = c->blocks[u >> 1];
      const int L = u & 1;
      const bool old_by_dev = b.dev_mapped[L] && b.dev_owned[L], new_by_pose = b.host_mapped[L] && b.pose_pending[L];
      if (old_by_dev || new_by_pose) {
        PoseUpd &pu = *pupd++;
        memset(&pu, 0, sizeof(pu));
        pu.block = u >> 1;
        pu.layer_flags = (uint32_t)L | (old_by_dev ? kPoseRemoveOld : 0u) | (new_by_pose ? 0u : kPoseNoInsert);
        pu.seq_local = b.map_seq[L];
        pu.root_flags = root_flags_of(c, b.model);
        pu.x = b.pose[L][0];
        pu.y = b.pose[L][1];
        pu.z = b.pose[L][2];
        pu.a = b.pose[L][3];
        if (old_by_dev)
          est_rem[L] += b.dev_steps[L];
        if (new_by_pose)
          est_ins[L] += b.host_steps[L];
      }
      if (b.dev_mapped[L] && !b.dev_owned[L]) // an outline the host rendered leaves through edge records, whatever replaces it
        rem_round[L] += emit_edges(b.dev_xy[L], u >> 1, L, rem_cursor, rem, &b.dev_poly[L]);
      if (new_by_pose) { // the device renders it: nothing of the outline is known here
        if (b.vec_used[L]) { // (the host-side pair was emptied when the pose was queued)
          b.dev_xy[L].clear();
          b.dev_poly[L].clear();
          b.vec_used[L] = false;
        }
        b.dev_steps[L] = b.host_steps[L];
        b.dev_mapped[L] = true;
        b.dev_owned[L] = true;
        b.pose_pending[L] = false;
        b.keep_seq[L] = false;
        b.dirty[L] = false;
        return;
      }
      b.dev_owned[L] = false;
      if (b.host_mapped[L]) {
        const uint32_t st = emit_edges(b.host_xy[L], u >> 1, L, ins_cursor, ins, &b.host_poly[L]);
        ins_round[L] += st;
        BlockUpd &up = *bupd++;
        up.block = u >> 1;
        up.model = b.model;
        up.zmin = b.zmin;
        up.zmax = b.zmax;
        up.seq_local = b.map_seq[L];
        up.layer = (uint32_t)L | (b.keep_seq[L] ? kUpdKeepSeq : 0u);
        up.root_flags = root_flags_of(c, b.model);
        up.pad = 0;
        b.keep_seq[L] = false;
        b.dev_xy[L].swap(b.host_xy[L]); // host_xy is rewritten by the next Map before it is read again
        b.dev_poly[L].swap(b.host_poly[L]);
        b.vec_used[L] = true;
        b.dev_steps[L] = b.host_steps[L];
      } else {
        b.dev_xy[L].clear();
        b.dev_poly[L].clear();
        b.dev_steps[L] = 0;
      }
      b.dev_mapped[L] = b.host_mapped[L];
      b.dirty[L] = false;
    };
    if (chunked && n_rem == 0 && n_ins == 0 && n_bupd == attr_blocks.size() && !c->pool.busy()) {
      // nothing but pose records (no unit emits an edge or a block update, so the cursors those share stay untouched):
      // stretch k's records start where the stretches before it end
      size_t first_pose[kUnitChunks + 1];
      first_pose[0] = 0;
      for (size_t k = 0; k < kUnitChunks; k++)
        first_pose[k + 1] = first_pose[k] + chunk_tally[k].n_pose;
      int64_t part_rem[kUnitChunks][2] = {}, part_ins[kUnitChunks][2] = {};
      const uint32_t *units = c->dirty_units.data() + unit;
      c->pool.parallel_for(kUnitChunks, 1, [&](size_t k0, size_t k1) {
        for (size_t k = k0; k < k1; k++) {
          PoseUpd *p = pupd + first_pose[k];
          BlockUpd *none = nullptr;
          uint32_t cur_r = 0, cur_i = 0;
          int64_t rr[2] = { 0, 0 }, ir[2] = { 0, 0 };
          for (size_t i = k * n_queued / kUnitChunks, e = (k + 1) * n_queued / kUnitChunks; i < e; i++)
            write_unit(units[i], p, none, cur_r, cur_i, part_rem[k], part_ins[k], rr, ir);
        }
      });
      for (size_t k = 0; k < kUnitChunks; k++)
        for (int L = 0; L < 2; L++) {
          est_rem[L] += part_rem[k][L];
          est_ins[L] += part_ins[k][L];
        }
      unit = end_unit;
    }
    for (; unit < end_unit; unit++)
      write_unit(c->dirty_units[unit], pupd, bupd, rem_cursor, ins_cursor, est_rem, est_ins, rem_round, ins_round);
    if (n_model)
      memcpy(mupd, c->model_upd.data(), n_model * sizeof(ModelUpd));
    memset(hdr, 0, sizeof(*hdr));
    hdr->n_remove = (uint32_t)n_rem;
    hdr->n_insert = (uint32_t)n_ins;
    hdr->n_block_upd = (uint32_t)n_bupd;
    hdr->n_model_upd = (uint32_t)n_model;
    hdr->steps_remove = rem_cursor;
    hdr->steps_insert = ins_cursor;
    hdr->rank = (uint32_t)rank;
    hdr->magic = kDeltaMagic;
    for (int L = 0; L < 2; L++) {
      hdr->steps_remove_l[L] = (uint32_t)rem_round[L];
      hdr->steps_insert_l[L] = (uint32_t)ins_round[L];
    }
    hdr->n_bytes = (uint32_t)bytes;
    hdr->n_pose_upd = (uint32_t)n_pose;
    for (int L = 0; L < 2; L++) {
      hdr->est_remove_l[L] = (uint32_t)est_rem[L];
      hdr->est_insert_l[L] = (uint32_t)est_ins[L];
      ins_round[L] += est_ins[L]; // slots the insertions may claim, for the compaction policy
    }
    hdr->epoch = c->epoch + 1; // what this sender would call the pass; a gathered pass takes the largest proposal
    const size_t nbytes = bytes;
    lap(3);
    CU(c, cudaMemcpyAsync(c->d_blob.p, c->h_blob.p, nbytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaEventRecord(c->blob_uploaded, c->stream));
    c->stats.h2d_bytes += nbytes;
    c->stats.edge_steps += rem_cursor + ins_cursor;
    if (!apply)
      for (int L = 0; L < 2; L++) {
        c->last_export_ins[L] = ins_round[L];
        c->last_export_rem[L] = rem_round[L] + est_rem[L];
      }
    for (int L = 0; L < 2; L++)
      c->live_entries[L] += d_live[L];
    if (apply) {
      if (!c->hold_epoch)
        c->epoch++;
      const int what = ((n_rem + n_ins + n_bupd + n_model) ? kDeltaEdgeRecords : 0) | (n_pose ? kDeltaPoseRecords : 0);
      int rc = summaries_fence_locked(c);
      if (rc)
        return rc;
      CU(c, cudaEventRecord(c->t_ev[0][0], c->stream));
      launch_apply_deltas(c->g, c->gv, (const unsigned char *)c->d_blob.p, 1, c->blob_cap, c->epoch, 0, what, c->stream,
                          &c->stats.launches_rasterise);
      if ((rc = ensure_table_room(c, ins_round))) // live entries grew by the net change, used slots by every insertion
        return rc;
      launch_apply_deltas(c->g, c->gv, (const unsigned char *)c->d_blob.p, 1, c->blob_cap, c->epoch, 1, what, c->stream,
                          &c->stats.launches_rasterise);
      if (owners_suspect)
        launch_owner_rebuild(c->g, c->stream, &c->stats.launches_rasterise);
      CU(c, cudaEventRecord(c->t_ev[0][1], c->stream));
      if ((rc = summaries_launch_locked(c)))
        return rc;
      c->t_rec[0] = true;
      CU(c, cudaGetLastError());
    }
    lap(4);
    if (out_bytes)
      *out_bytes = nbytes;
    first_round = false;
    if (unit < c->dirty_units.size()) // another round follows and reuses the pinned blob right away
      CU(c, cudaStreamSynchronize(c->stream));
  }
  c->dirty_units.clear();
  c->model_upd.clear();
  c->local_seq = 0;
  return 0;
}

int apply_deltas_locked(stg_rt_ctx *c)
{
  if (c->dirty_units.empty() && c->model_upd.empty())
    return 0;
  return build_and_send_deltas(c, true, 0, nullptr);
}

// ---- fans -------------------------------------------------------------------------------------

uint32_t want_of(const stg_rt_fan_out &o)
{
  return (o.ranges ? STG_RT_WANT_RANGES : 0) | (o.intensities ? STG_RT_WANT_INTENSITIES : 0) |
         (o.bearings ? STG_RT_WANT_BEARINGS : 0) | (o.hit_model ? STG_RT_WANT_HIT_MODEL : 0) |
         (o.hit_cell ? STG_RT_WANT_HIT_CELL : 0) | (o.steps ? STG_RT_WANT_STEPS : 0);
}

// A K1 pass must not start while k1_summaries of the pass before is still reading the tiles ...
int summaries_fence_locked(stg_rt_ctx *c)
{
  if (c->sum_pending)
    CU(c, cudaStreamWaitEvent(c->stream, c->sum_done, 0));
  return 0;
}
// ... and hands the regions it took bits out of to the next one (behind the pass, beside the main stream's next work)
int summaries_launch_locked(stg_rt_ctx *c)
{
  CU(c, cudaEventRecord(c->sum_ready, c->stream));
  CU(c, cudaStreamWaitEvent(c->sum_stream, c->sum_ready, 0));
  launch_summaries(c->g, c->sum_stream);
  CU(c, cudaEventRecord(c->sum_done, c->sum_stream));
  c->sum_pending = true;
  c->stats.launches_rasterise++;
  return 0;
}

int reserve_outputs(stg_rt_ctx *c, FanArrays &a, uint64_t beams, uint32_t want)
{
  int rc = dev_reserve(c, a.heading, beams * 8);
  if (!rc && (want & STG_RT_WANT_RANGES)) rc = dev_reserve(c, a.ranges, beams * 8);
  if (!rc && (want & STG_RT_WANT_INTENSITIES)) rc = dev_reserve(c, a.intens, beams * 8);
  if (!rc && (want & STG_RT_WANT_BEARINGS)) rc = dev_reserve(c, a.bearings, beams * 8);
  if (!rc && (want & STG_RT_WANT_HIT_MODEL)) rc = dev_reserve(c, a.hit_model, beams * 4);
  if (!rc && (want & STG_RT_WANT_HIT_CELL)) rc = dev_reserve(c, a.hit_cell, beams * 8);
  if (!rc && (want & STG_RT_WANT_STEPS)) rc = dev_reserve(c, a.steps, beams * 8);
  return rc;
}

uint32_t uniform_samples_of(const FanRec *fans, size_t n)
{
  for (size_t i = 1; i < n; i++)
    if (fans[i].n_samples != fans[0].n_samples)
      return 0;
  return n ? fans[0].n_samples : 0;
}

FanBatchView view_of(const FanArrays &a, uint32_t n_fans, uint64_t beams, uint32_t want, bool trig, bool headings)
{
  FanBatchView v;
  memset(&v, 0, sizeof(v));
  v.n_fans = n_fans;
  v.n_beams = beams;
  v.beam_begin = 0;
  v.beam_end = beams;
  v.noise = nullptr;
  v.noise_epoch = 0;
  v.fans = (const FanRec *)a.fans.p;
  v.fan_first_beam = (const uint32_t *)a.first.p;
  v.headings_in = headings ? (const double *)a.headings_in.p : nullptr;
  v.trig_in = trig ? (const double *)a.trig.p : nullptr;
  v.heading = (double *)a.heading.p;
  v.ranges = (want & STG_RT_WANT_RANGES) ? (double *)a.ranges.p : nullptr;
  v.intensities = (want & STG_RT_WANT_INTENSITIES) ? (double *)a.intens.p : nullptr;
  v.bearings = (want & STG_RT_WANT_BEARINGS) ? (double *)a.bearings.p : nullptr;
  v.hit_model = (want & STG_RT_WANT_HIT_MODEL) ? (int32_t *)a.hit_model.p : nullptr;
  v.hit_cell = (want & STG_RT_WANT_HIT_CELL) ? (int32_t *)a.hit_cell.p : nullptr;
  v.steps = (want & STG_RT_WANT_STEPS) ? (uint32_t *)a.steps.p : nullptr;
  return v;
}

int pinfans_resize(stg_rt_ctx *c, size_t n)
{
  stg_rt_ctx::PinFans &q = c->q_fans;
  if (n > q.cap) {
    const size_t cap = std::max<size_t>(std::max(n, 2 * q.cap), 1024);
    FanRec *p = nullptr;
    if (cudaMallocHost(&p, cap * sizeof(FanRec)) != cudaSuccess)
      return fail(c, STG_RT_ENOMEM, "cudaMallocHost(%zu) for the fan queue failed", cap * sizeof(FanRec));
    if (q.n)
      memcpy(p, q.p, q.n * sizeof(FanRec));
    if (q.p)
      cudaFreeHost(q.p);
    q.p = p;
    q.cap = cap;
  }
  q.n = n;
  return 0;
}

int validate_fans(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const double *headings)
{
  for (uint32_t i = 0; i < n_fans; i++) {
    const stg_rt_fan &f = fans[i];
    if (f.finder >= c->cfg.max_models || !c->model_known[f.finder])
      return fail(c, STG_RT_EINVAL, "fan %u: unknown finder model %u", i, f.finder);
    if (f.pred > STG_RT_PRED_GRIPPER || f.angles > STG_RT_ANGLES_EXPLICIT || f.layer > 1)
      return fail(c, STG_RT_EINVAL, "fan %u: bad pred/angles/layer", i);
    if (f.angles == STG_RT_ANGLES_EXPLICIT && !headings)
      return fail(c, STG_RT_EINVAL, "fan %u: explicit headings missing", i);
  }
  return 0;
}

// copy one result array back: straight into caller memory when it is page-locked memory of ours,
// otherwise into the pinned staging buffer (delivered by deliver_locked)
// [lo, hi) = the beams (batch-wide indices) to fetch now; a submit covers [first, first + n)
int fetch_array(stg_rt_ctx *c, void *user, const void *dev, PinBuf &stage, uint64_t first, uint64_t n, size_t elt,
                uint64_t lo, uint64_t hi, cudaStream_t on)
{
  if (!user)
    return 0;
  lo = std::max(lo, first);
  hi = std::min(hi, first + n);
  if (hi <= lo)
    return 0;
  const size_t bytes = (hi - lo) * elt;
  void *dst = is_host_alloc(c, user, n * elt) ? (char *)user + (lo - first) * elt : (char *)stage.p + lo * elt;
  CU(c, cudaMemcpyAsync(dst, (const char *)dev + lo * elt, bytes, cudaMemcpyDeviceToHost, on));
  c->stats.d2h_bytes += bytes;
  return 0;
}

void deliver_array(stg_rt_ctx *c, void *user, const PinBuf &stage, uint64_t first, uint64_t n, size_t elt, bool direct)
{
  if (!user || direct)
    return;
  memcpy(user, (const char *)stage.p + first * elt, n * elt);
}

int deliver_post_locked(stg_rt_ctx *c)
{
  if (!c->fid_job.pending && !c->blob_job.pending && !c->col_job.pending && !c->touch_job.pending)
    return 0;
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->touch_job.pending) {
    const stg_rt_ctx::ToucherJob &j = c->touch_job;
    const uint32_t *cnt = (const uint32_t *)c->h_touch_out.p, *ids = cnt + j.n_queries;
    memcpy(j.user_count, cnt, (size_t)j.n_queries * 4);
    for (uint32_t q = 0; q < j.n_queries; q++)
      memcpy(j.user_out + (size_t)q * j.max_per_query, ids + (size_t)q * j.max_per_query,
             (size_t)std::min(cnt[q] == 0xffffffffu ? j.max_per_query : cnt[q], j.max_per_query) * 4);
    c->touch_job.pending = false;
  }
  if (c->col_job.pending) {
    const int32_t *hit = (const int32_t *)c->h_col_hit.p;
    for (size_t q = 0; q < c->col_job.fallback.size(); q++)
      c->col_job.user_hit[q] = hit[q] >= 0 ? hit[q] : c->col_job.fallback[q];
    c->col_job.pending = false;
  }
  if (c->fid_job.pending) { // unpack: finder f's records follow finder f-1's in the staging buffer
    const double t_prof = prof_now(c);
    const uint32_t *cnt = (const uint32_t *)c->h_fid_count.p;
    if (c->cfg.flags & STG_RT_CFG_HOST_EXACT_FIDUCIALS) {
      // range and bearing once more with the HOST C library's hypot / atan2 (model_fiducial.cc:130,144), from the
      // records the submit staged: a caller that must agree with a CPU run to the last bit (the libstage binding)
      // gets the reference's own arithmetic; what was seen was decided on the device
      const char *in = (const char *)c->h_fid_in.p;
      const FidTargetRec *tg = (const FidTargetRec *)in;
      const FidFinderRec *fdr = (const FidFinderRec *)(in + (size_t)c->fid_job.n_targets * sizeof(FidTargetRec));
      FiducialRec *rec = (FiducialRec *)c->h_fid_out.p;
      auto norm = [](double a) {
        while (a < -M_PI)
          a += 2.0 * M_PI;
        while (a > M_PI)
          a -= 2.0 * M_PI;
        return a;
      };
      for (uint32_t f = 0; f < c->fid_job.n_finders; f++) {
        const FidFinderRec &fd = fdr[f];
        for (uint32_t k = 0, n = std::min(cnt[f], c->fid_job.max_per_finder); k < n; k++, rec++) {
          const FidTargetRec &him = tg[rec->target];
          const double dx = him.x - fd.x, dy = him.y - fd.y;
          rec->range = hypot(dy, dx);
          rec->bearing = norm(atan2(dy, dx) - fd.a);
          rec->id = rec->range < fd.max_range_id ? him.fiducial_return : 0;
        }
      }
    }
    const FiducialRec *packed = (const FiducialRec *)c->h_fid_out.p;
    FiducialRec *user = (FiducialRec *)c->fid_job.user_out;
    size_t at = 0;
    if (c->fid_job.packed_capacity) { // the caller takes them as they are: finder f's records follow finder f-1's
      for (uint32_t f = 0; f < c->fid_job.n_finders; f++)
        at += std::min(cnt[f], c->fid_job.max_per_finder);
      // (config 4: 65 MB of records per tick - 4.5 ms for one core, so the host threads share the copy; a streamed
      // expansion still open on them belongs to a march that ended before these records were made: it is closed first)
      if (c->pool.busy())
        c->pool.finish();
      const size_t bytes = std::min<size_t>(at, c->fid_job.packed_capacity) * sizeof(FiducialRec);
      char *dst = (char *)user;
      const char *src = (const char *)packed;
      c->pool.parallel_for(bytes, (size_t)2 << 20, [dst, src](size_t lo, size_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
    } else {
      for (uint32_t f = 0; f < c->fid_job.n_finders; f++) {
        const uint32_t k = std::min(cnt[f], c->fid_job.max_per_finder);
        if (k)
          memcpy(user + (size_t)f * c->fid_job.max_per_finder, packed + at, (size_t)k * sizeof(FiducialRec));
        at += k;
      }
    }
    memcpy(c->fid_job.user_count, cnt, (size_t)c->fid_job.n_finders * 4);
    c->stats.d2h_bytes += at * sizeof(FiducialRec) + (size_t)c->fid_job.n_finders * 4;
    c->fid_job.pending = false;
    if (c->prof_on)
      c->prof_us[7] += prof_now(c) - t_prof;
  }
  if (c->blob_job.pending) {
    memcpy(c->blob_job.user_out, c->h_blob_out.p, c->blob_job.out_bytes);
    memcpy(c->blob_job.user_count, c->h_blob_count.p, c->blob_job.count_bytes);
    c->blob_job.pending = false;
  }
  return 0;
}

int deliver_locked(stg_rt_ctx *c)
{
  int prc = deliver_post_locked(c);
  if (prc)
    return prc;
  if (c->f_subs.empty())
    return 0;
  const double t_wait0 = prof_now(c);
  CU(c, cudaStreamSynchronize(c->stream));
  const double t_wait1 = prof_now(c);
  if (c->prof_on)
    c->prof_us[8] += t_wait1 - t_wait0;
  for (const Submit &s : c->f_subs) {
    // the kernel wrote ranges / intensities / hit models of a solo pinned submit in place
    deliver_array(c, s.out.ranges, c->h_ranges, s.first_beam, s.n_beams, 8,
                  c->f_copied ? is_host_alloc(c, s.out.ranges, s.n_beams * 8) : c->f_direct[0]);
    if (c->f_streaming && s.out.intensities) { // expanded while the kernel ran; whatever is left is finished here
      c->pool.finish();
      if (c->stream_gave_up.load()) { // (only after a failed launch) do it the plain way
        c->idx_clean_bytes = 0;
        const uint8_t *idx = (const uint8_t *)c->h_intens_idx.p + s.first_beam;
        for (uint64_t i = 0; i < s.n_beams; i++)
          s.out.intensities[i] = c->palette[idx[i]];
      }
    } else if (c->f_compact_intens && s.out.intensities) { // palette indices -> the caller's doubles (model_ranger.cc:250)
      c->idx_clean_bytes = 0;
      const uint8_t *idx = (const uint8_t *)c->h_intens_idx.p + s.first_beam;
      double *dst = s.out.intensities;
      const double *pal = c->palette.data();
      c->pool.parallel_for((size_t)s.n_beams, 1 << 16, [idx, dst, pal](size_t lo, size_t hi) {
        // streaming stores: the lines are written whole and not read again soon, so they need not be fetched first
        size_t i = lo;
        for (; i < hi && ((uintptr_t)(dst + i) & 15u); i++)
          dst[i] = pal[idx[i]];
        for (; i + 2 <= hi; i += 2)
          _mm_stream_pd(dst + i, _mm_set_pd(pal[idx[i + 1]], pal[idx[i]]));
        for (; i < hi; i++)
          dst[i] = pal[idx[i]];
        _mm_sfence();
      });
    } else {
      deliver_array(c, s.out.intensities, c->h_intens, s.first_beam, s.n_beams, 8,
                    c->f_copied ? is_host_alloc(c, s.out.intensities, s.n_beams * 8) : c->f_direct[1]);
    }
    deliver_array(c, s.out.hit_model, c->h_hit_model, s.first_beam, s.n_beams, 4,
                  c->f_copied ? is_host_alloc(c, s.out.hit_model, s.n_beams * 4) : c->f_direct[2]);
    // these were copied device-to-host, straight into pinned caller memory when possible
    deliver_array(c, s.out.bearings, c->h_bearings, s.first_beam, s.n_beams, 8, is_host_alloc(c, s.out.bearings, s.n_beams * 8));
    deliver_array(c, s.out.hit_cell, c->h_hit_cell, s.first_beam, s.n_beams, 8, is_host_alloc(c, s.out.hit_cell, s.n_beams * 8));
    deliver_array(c, s.out.steps, c->h_steps, s.first_beam, s.n_beams, 8, is_host_alloc(c, s.out.steps, s.n_beams * 8));
  }
  if (c->prof_on)
    c->prof_us[9] += prof_now(c) - t_wait1;
  c->f_subs.clear();
  c->f_beams = 0;
  return 0;
}

int apply_deltas_locked(stg_rt_ctx *c);

// with_deltas: apply the queued deltas between the headings kernel (which does not read the grid, so
// it can run while the host is still building the delta blob) and the march (which must see them).
int launch_fans_locked(stg_rt_ctx *c, bool with_deltas = false)
{
  if (c->q_subs.empty())
    return with_deltas ? apply_deltas_locked(c) : 0;
  const double t_prof0 = prof_now(c);
  int rc = deliver_locked(c); // staging buffers are single-buffered
  if (rc)
    return rc;
  const bool rig = c->q_rig.active;
  const uint32_t n_fans = rig ? c->q_rig.n_rangers * c->q_rig.n_sensors : (uint32_t)c->q_fans.size();
  const uint64_t beams = c->q_beams;
  uint32_t want = 0;
  for (const Submit &s : c->q_subs)
    want |= want_of(s.out);
  if (!rig)
    c->q_first.push_back((uint32_t)beams);

  // one pinned block: fans | first | headings | trig
  const size_t b_fans = (size_t)n_fans * sizeof(FanRec), b_first = (size_t)(n_fans + 1) * 4;
  const size_t b_head = c->q_headings.size() * 8, b_trig = c->q_has_trig ? beams * 16 : 0;
  const size_t o_first = b_fans, o_head = (o_first + b_first + 15) & ~(size_t)15, o_trig = o_head + b_head;
  const size_t b_noise = c->q_has_noise ? (size_t)n_fans * sizeof(FanNoiseRec) : 0, o_noise = o_trig + b_trig, b_all = o_noise + b_noise;
  if (!rig && (rc = pin_reserve(c, c->h_in, b_all))) // a ranger-model batch has nothing to stage here
    return rc;
  char *hp = (char *)c->h_in.p;
  if (!rig)
    memcpy(hp + o_first, c->q_first.data(), b_first);
  if (b_head)
    memcpy(hp + o_head, c->q_headings.data(), b_head);
  if (b_trig)
    memcpy(hp + o_trig, c->q_trig.data(), b_trig);
  if (b_noise) {
    c->q_noise.resize(n_fans); // fans queued without noise after the last noisy one: zeros
    memcpy(hp + o_noise, c->q_noise.data(), b_noise);
  }
  // one device block with the same layout as the pinned one: a single upload per batch
  if ((rc = dev_reserve(c, c->d_in, b_all + 16)) || (rc = reserve_outputs(c, c->d, beams, want)))
    return rc;
  if (rig) { // fan records and first-beam prefix are derived on the device (k2_expand_rangers)
    if ((rc = dev_reserve(c, c->d_rig, c->q_rig.bytes)))
      return rc;
    CU(c, cudaMemcpyAsync(c->d_rig.p, c->h_rig.p, c->q_rig.bytes, cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += c->q_rig.bytes;
    const char *dr = (const char *)c->d_rig.p;
    launch_expand_rangers((const RangerSensorRec *)dr, c->q_rig.n_sensors, (const uint32_t *)(dr + c->q_rig.o_first),
                          (const RangerPoseRec *)(dr + c->q_rig.o_poses), c->q_rig.n_rangers, (FanRec *)c->d_in.p,
                          (uint32_t *)((char *)c->d_in.p + o_first), c->stream);
    c->stats.launches_post++;
  } else {
    CU(c, cudaMemcpyAsync(c->d_in.p, c->q_fans.data(), b_fans, cudaMemcpyHostToDevice, c->stream)); // from the queue itself
    CU(c, cudaMemcpyAsync((char *)c->d_in.p + o_first, hp + o_first, b_all - o_first, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaEventRecord(c->fans_uploaded, c->stream));
    c->stats.h2d_bytes += b_all;
  }

  if ((want & STG_RT_WANT_RANGES) && (rc = pin_reserve(c, c->h_ranges, beams * 8))) return rc;
  if ((want & STG_RT_WANT_INTENSITIES) && (rc = pin_reserve(c, c->h_intens, beams * 8))) return rc;
  if ((want & STG_RT_WANT_BEARINGS) && (rc = pin_reserve(c, c->h_bearings, beams * 8))) return rc;
  if ((want & STG_RT_WANT_HIT_MODEL) && (rc = pin_reserve(c, c->h_hit_model, beams * 4))) return rc;
  if ((want & STG_RT_WANT_HIT_CELL) && (rc = pin_reserve(c, c->h_hit_cell, beams * 8))) return rc;
  if ((want & STG_RT_WANT_STEPS) && (rc = pin_reserve(c, c->h_steps, beams * 8))) return rc;

  FanBatchView v = view_of(c->d, n_fans, beams, want, c->q_has_trig, b_head != 0);
  v.uniform_samples = rig ? c->q_rig.uniform_samples : uniform_samples_of(c->q_fans.data(), n_fans);
  v.fans = (const FanRec *)c->d_in.p;
  v.fan_first_beam = (const uint32_t *)((const char *)c->d_in.p + o_first);
  v.headings_in = b_head ? (const double *)((const char *)c->d_in.p + o_head) : nullptr;
  v.trig_in = c->q_has_trig ? (const double *)((const char *)c->d_in.p + o_trig) : nullptr;
  v.noise = b_noise ? (const FanNoiseRec *)((const char *)c->d_in.p + o_noise) : nullptr;
  v.noise_epoch = ++c->noise_epoch;
  // Results go straight to page-locked host memory from inside the march kernel (zero-copy stores
  // over PCIe, overlapped with the marching of the other warps) instead of a device buffer plus a
  // device-to-host copy afterwards: into the caller's own arrays when the batch is one submit whose
  // arrays came from stg_rt_host_alloc, otherwise into the pinned staging arrays that stg_rt_sync
  // copies from. Diagnostics (hit cell, step counters) and bearings keep the device buffer + copy.
  const Submit *solo = c->q_subs.size() == 1 ? &c->q_subs[0] : nullptr;
  auto host_target = [&](void *user, PinBuf &stage, size_t elt, bool &direct) -> void * {
    direct = solo && user && is_host_alloc(c, user, beams * elt);
    return direct ? user : stage.p;
  };
  c->f_direct[0] = c->f_direct[1] = c->f_direct[2] = false;
  bool zc_ranges = false, zc_intens = false, zc_model = false;
  // STG_RT_RESULTS=copy: device buffers + copy engine instead of zero-copy stores (measurements)
  static const bool copy_results = getenv("STG_RT_RESULTS") && !strcmp(getenv("STG_RT_RESULTS"), "copy");
  const bool want_intens_copy = copy_results && (want & STG_RT_WANT_INTENSITIES);
  (void)want_intens_copy;
  if (copy_results)
    want &= ~(STG_RT_WANT_RANGES | STG_RT_WANT_INTENSITIES | STG_RT_WANT_HIT_MODEL); // for the zero-copy set-up below only
  if (want & STG_RT_WANT_RANGES) {
    v.ranges = (double *)host_target(solo ? solo->out.ranges : nullptr, c->h_ranges, 8, c->f_direct[0]);
    zc_ranges = true;
  }
  c->f_compact_intens = false;
  static const bool wide_intens = getenv("STG_RT_INTENSITIES") && !strcmp(getenv("STG_RT_INTENSITIES"), "f64"); // measurements
  if ((want & STG_RT_WANT_INTENSITIES) && c->palette_ok && !wide_intens) {
    if ((rc = pin_reserve(c, c->h_intens_idx, beams)))
      return rc;
    v.intensities = nullptr;
    v.intens_idx = (uint8_t *)c->h_intens_idx.p;
    c->f_compact_intens = true;
  } else if (want & STG_RT_WANT_INTENSITIES) {
    v.intensities = (double *)host_target(solo ? solo->out.intensities : nullptr, c->h_intens, 8, c->f_direct[1]);
    zc_intens = true;
  }
  if (want & STG_RT_WANT_HIT_MODEL) {
    v.hit_model = (int32_t *)host_target(solo ? solo->out.hit_model : nullptr, c->h_hit_model, 4, c->f_direct[2]);
    zc_model = true;
  }
  // The headings do not read the grid: when a delta pass is about to run they go on the side stream and overlap it
  // (the march waits for both).
  const bool side = with_deltas && (!c->dirty_units.empty() || !c->model_upd.empty());
  cudaStream_t hs = side ? c->copy_stream : c->stream;
  if (side) {
    CU(c, cudaEventRecord(c->chunk_ev[0], c->stream)); // the fan records are uploaded
    CU(c, cudaStreamWaitEvent(hs, c->chunk_ev[0], 0));
  }
  const bool refill = ray_march_refill_wanted(v);
  if (refill && !v.trig_in && (rc = dev_reserve(c, c->d.trigw, beams * 16)))
    return rc;
  CU(c, cudaEventRecord(c->t_ev[1][0], hs));
  launch_fan_headings(v, hs);
  if (refill && !v.trig_in) { // cos / sin of every heading with all lanes busy, instead of inside the march's refills
    launch_beam_trig(v, (double *)c->d.trigw.p, hs);
    c->stats.launches_post++;
  }
  CU(c, cudaEventRecord(c->t_ev[1][1], hs));
  const double t_prof1 = prof_now(c);
  if (with_deltas && (rc = apply_deltas_locked(c)))
    return rc;
  if (side) {
    CU(c, cudaEventRecord(c->chunk_ev[1], hs));
    CU(c, cudaStreamWaitEvent(c->stream, c->chunk_ev[1], 0));
  }
  const double t_prof2 = prof_now(c);
  // A large batch of one submit that wants intensities as doubles: stream them (see h_done_flags), in stretches of 8192
  // beams (256 warps): a stretch is expanded as soon as the march has reported it complete
  constexpr uint32_t kStretchShift = 13;
  constexpr size_t kStretch = (size_t)1 << kStretchShift;
  const char *stream_env = getenv("STG_RT_STREAM");
  const bool no_stream = stream_env && !strcmp(stream_env, "0");
  c->f_streaming = false;
  if (c->f_compact_intens && solo && solo->out.intensities && !refill && !no_stream && beams >= (1u << 17) && !c->pool.busy()) {
    const size_t n_flags = (size_t)((beams + kStretch - 1) / kStretch);
    const bool fresh = n_flags * 4 > c->h_done_flags.cap;
    if ((rc = pin_reserve(c, c->h_done_flags, n_flags * 4)) || (rc = dev_reserve(c, c->d_stretch_count, c->h_done_flags.cap)))
      return rc;
    if (fresh) { // (every launch leaves the counters at zero again)
      CU(c, cudaMemsetAsync(c->d_stretch_count.p, 0, c->d_stretch_count.cap, c->stream));
      memset(c->h_done_flags.p, 0, c->h_done_flags.cap);
      c->done_seq = 0;
    }
    if (++c->done_seq == 0) { // wrapped: start over with clean flags
      memset(c->h_done_flags.p, 0, c->h_done_flags.cap);
      c->done_seq = 1;
    }
    v.stretch_count = (uint32_t *)c->d_stretch_count.p;
    v.stretch_done = (uint32_t *)c->h_done_flags.p;
    v.done_seq = c->done_seq;
    v.stretch_shift = kStretchShift;
    c->f_streaming = true;
    // every staging byte the march is about to write holds "not written yet" (normally left so by the last expansion)
    if (c->idx_clean_base != c->h_intens_idx.p || c->idx_clean_bytes < beams) {
      memset(c->h_intens_idx.p, (int)kPaletteUnwritten, (size_t)beams);
      c->idx_clean_base = c->h_intens_idx.p;
      c->idx_clean_bytes = (size_t)beams;
    }
  } else if (c->f_compact_intens) {
    c->idx_clean_bytes = 0; // the march writes the staging bytes and nobody puts them back
  }
  CU(c, cudaEventRecord(c->t_ev[2][0], c->stream));
  if (refill)
    launch_ray_march_refill(c->g, v, v.trig_in ? v.trig_in : (const double *)c->d.trigw.p, (uint32_t *)c->d_work_counter.p, c->stream);
  else
    launch_ray_march(c->g, v, c->stream);
  CU(c, cudaEventRecord(c->t_ev[2][1], c->stream));
  c->t_rec[1] = c->t_rec[2] = true;
  CU(c, cudaGetLastError());
  if (c->f_streaming) {
    uint8_t *idx = (uint8_t *)c->h_intens_idx.p;
    double *dst = solo->out.intensities;
    const double *pal = c->palette.data();
    const volatile uint32_t *flags = (const volatile uint32_t *)c->h_done_flags.p;
    const uint32_t seq = c->done_seq;
    const size_t n = (size_t)beams;
    std::atomic<bool> *gave_up = &c->stream_gave_up;
    gave_up->store(false);
    c->pool.begin((n + kStretch - 1) / kStretch, [idx, dst, pal, flags, seq, n, gave_up](size_t part) {
      const size_t lo = part * kStretch, hi = std::min(n, lo + kStretch);
      const std::chrono::steady_clock::time_point until = std::chrono::steady_clock::now() + std::chrono::seconds(2);
      auto late = [&]() { // the kernel did not finish (an error the sync will report): nothing to expand
        if (gave_up->load(std::memory_order_relaxed) || std::chrono::steady_clock::now() > until) {
          gave_up->store(true);
          return true;
        }
        _mm_pause();
        return false;
      };
      while (flags[part] != seq)
        if (late())
          return;
      std::atomic_thread_fence(std::memory_order_acquire);
      // belt and braces: a byte that still reads "not written yet" is a store that has not arrived; wait for it
      for (const uint8_t *q = (const uint8_t *)memchr(idx + lo, (int)kPaletteUnwritten, hi - lo); q;
           q = (const uint8_t *)memchr(q, (int)kPaletteUnwritten, (size_t)(idx + hi - q))) {
        while (*(const volatile uint8_t *)q == kPaletteUnwritten)
          if (late())
            return;
      }
      size_t i = lo;
      for (; i < hi && ((uintptr_t)(dst + i) & 15u); i++)
        dst[i] = pal[idx[i]];
      for (; i + 2 <= hi; i += 2)
        _mm_stream_pd(dst + i, _mm_set_pd(pal[idx[i + 1]], pal[idx[i]]));
      for (; i < hi; i++)
        dst[i] = pal[idx[i]];
      _mm_sfence();
      memset(idx + lo, (int)kPaletteUnwritten, hi - lo); // read: ready for the next tick's march
    });
  }
  c->stats.launches_march += 1;
  c->stats.launches_post += 1;
  c->stats.beams += beams;
  c->stats.fans += n_fans;
  c->stats.d2h_bytes += beams * ((zc_ranges ? 8 : 0) + (zc_intens ? 8 : 0) + (zc_model ? 4 : 0) + (c->f_compact_intens ? 1 : 0));
  for (const Submit &s : c->q_subs) { // what did not go zero-copy
    if (copy_results) {
      if ((rc = fetch_array(c, s.out.ranges, c->d.ranges.p, c->h_ranges, s.first_beam, s.n_beams, 8, 0, beams, c->stream)) ||
          (rc = fetch_array(c, s.out.intensities, c->d.intens.p, c->h_intens, s.first_beam, s.n_beams, 8, 0, beams, c->stream)) ||
          (rc = fetch_array(c, s.out.hit_model, c->d.hit_model.p, c->h_hit_model, s.first_beam, s.n_beams, 4, 0, beams, c->stream)))
        return rc;
    }
    if ((rc = fetch_array(c, s.out.bearings, c->d.bearings.p, c->h_bearings, s.first_beam, s.n_beams, 8, 0, beams, c->stream)) ||
        (rc = fetch_array(c, s.out.hit_cell, c->d.hit_cell.p, c->h_hit_cell, s.first_beam, s.n_beams, 8, 0, beams, c->stream)) ||
        (rc = fetch_array(c, s.out.steps, c->d.steps.p, c->h_steps, s.first_beam, s.n_beams, 8, 0, beams, c->stream)))
      return rc;
  }

  c->f_copied = copy_results;
  c->f_subs.swap(c->q_subs);
  c->f_beams = beams;
  c->q_subs.clear();
  c->q_fans.clear();
  c->q_rig.active = false;
  c->q_first.clear();
  c->q_headings.clear();
  c->q_trig.clear();
  c->q_noise.clear();
  c->q_has_noise = false;
  c->q_beams = 0;
  c->q_has_trig = false;
  if (c->prof_on) {
    c->prof_us[0] += t_prof1 - t_prof0;
    c->prof_us[5] += prof_now(c) - t_prof2;
    c->prof_flushes++;
  }
  return 0;
}

int flush_locked(stg_rt_ctx *c)
{
  return launch_fans_locked(c, true);
}

// What K1 could not do and said so (GridView::err_host); the stream must be idle. The grid may be incomplete after
// any of these: the caller should treat the context as lost (libstage falls back to its CPU path).
int device_errors_locked(stg_rt_ctx *c)
{
  const uint32_t e = *(volatile uint32_t *)c->err_host;
  if (!e)
    return STG_RT_OK;
  *(volatile uint32_t *)c->err_host = 0;
  if (e & kErrTableFull)
    return fail(c, STG_RT_ENOMEM, "the device ran out of cell-entry slots while inserting (raise max_cell_entries); entries were dropped");
  if (e & kErrBadBlob)
    return fail(c, STG_RT_EINVAL, "a delta slot did not hold a well-formed blob (gathered before the export had landed, or a foreign "
                                  "buffer); it was skipped");
  return fail(c, STG_RT_EINVAL, "a delta record named a cell outside the device extent or an unknown block; it was skipped");
}

void touch_unit(stg_rt_ctx *c, uint32_t block, int layer)
{
  BlockShadow &b = c->blocks[block];
  const uint32_t key = b.host_mapped[layer] ? b.map_seq[layer] : 0u; // sort_units' key, as the caller has just set it
  if (!b.dirty[layer]) {
    b.dirty[layer] = true;
    if (c->dirty_units.empty())
      c->units_in_order = true;
    else if (key < c->last_unit_key)
      c->units_in_order = false;
    c->last_unit_key = key;
    c->dirty_units.push_back(block * 2 + (uint32_t)layer);
  } else {
    c->units_in_order = false; // a queued unit's key changed where it stands
  }
}

} // namespace

// ================================================================================================
// C ABI
// ================================================================================================

extern "C" {

int stg_rt_abi_version(void) { return STG_RT_ABI_VERSION; }

const char *stg_rt_last_error(const stg_rt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int stg_rt_create(const stg_rt_config *cfg, stg_rt_ctx **out)
{
  if (!cfg || !out || cfg->struct_size != sizeof(stg_rt_config))
    return fail(nullptr, STG_RT_EINVAL, "stg_rt_create: bad config (struct_size %u, expected %zu)",
                cfg ? cfg->struct_size : 0u, sizeof(stg_rt_config));
  if (!(cfg->ppm > 0) || cfg->sreg_nx < 1 || cfg->sreg_ny < 1 || (int64_t)cfg->sreg_nx * cfg->sreg_ny > 4095 ||
      cfg->max_models < 1 || cfg->max_models > kRecRootMask || cfg->max_blocks < 1)
    return fail(nullptr, STG_RT_EINVAL, "stg_rt_create: bad ppm / extent / capacities");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
    return fail(nullptr, STG_RT_ENODEV, "stg_rt_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(cudaGetLastError()));
  if (cfg->device < 0 || cfg->device >= n_dev)
    return fail(nullptr, STG_RT_ENODEV, "stg_rt_create: device %d of %d", cfg->device, n_dev);
  stg_rt_ctx *c = new stg_rt_ctx;
  c->cfg = *cfg;
#define CUC(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      fail(nullptr, STG_RT_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));                        \
      delete c;                                                                                    \
      return STG_RT_ECUDA;                                                                         \
    }                                                                                              \
  } while (0)
  CUC(cudaSetDevice(cfg->device));
  CUC(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CUC(cudaEventCreateWithFlags(&c->blob_uploaded, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&c->fans_uploaded, cudaEventDisableTiming));
  for (int k = 0; k < 6; k++)
    for (int j = 0; j < 2; j++)
      CUC(cudaEventCreate(&c->t_ev[k][j]));
  CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&c->sum_stream, cudaStreamNonBlocking));
  CUC(cudaEventCreateWithFlags(&c->sum_ready, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&c->sum_done, cudaEventDisableTiming));
  for (int k = 0; k < 8; k++)
    CUC(cudaEventCreateWithFlags(&c->chunk_ev[k], cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&c->copies_done, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&c->hdrs_ready, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&c->xstream_ev, cudaEventDisableTiming));
  GridView &g = c->g;
  g.ppm = cfg->ppm;
  g.rx0 = cfg->sreg_x0 * 32;
  g.ry0 = cfg->sreg_y0 * 32;
  g.nrx = cfg->sreg_nx * 32;
  g.nry = cfg->sreg_ny * 32;
  c->n_regions = (size_t)g.nrx * g.nry;
  c->cell_x0 = g.rx0 * 32;
  c->cell_y0 = g.ry0 * 32;
  c->cell_x1 = c->cell_x0 + g.nrx * 32 - 1;
  c->cell_y1 = c->cell_y0 + g.nry * 32 - 1;
  const uint64_t entries = cfg->max_cell_entries ? cfg->max_cell_entries : (1u << 22);
  c->slot_cap = std::max<uint32_t>(next_pow2(2 * entries), 16 * kSectorSlots);
  g.slot_mask = c->slot_cap - 1;
  g.n_blocks = cfg->max_blocks;
  g.n_models = cfg->max_models;
  CUC(cudaMalloc(&g.head, (c->n_regions + 1) * sizeof(RegionHead))); // + the record that holds the "owners are stale" flag
  CUC(cudaMemsetAsync(g.head, 0, (c->n_regions + 1) * sizeof(RegionHead), c->stream));
  g.n_regions = (uint32_t)c->n_regions;
  // both layers' tiles in one allocation: the march addresses them with one 32-bit word index
  // (and behind them both layers' summary words)
  CUC(cudaMalloc(&g.occ[0], occ_alloc_words(c->n_regions) * 4));
  CUC(cudaMemsetAsync(g.occ[0], 0, occ_alloc_words(c->n_regions) * 4, c->stream));
  g.occ[1] = g.occ[0] + c->n_regions * kTileWords;
  g.occ_dirty = g.occ[0] + 2 * c->n_regions * kTileWords;
  for (int L = 0; L < 2; L++) {
    CUC(cudaMalloc(&g.slots[L], (size_t)c->slot_cap * 8));
    CUC(cudaMemsetAsync(g.slots[L], 0, (size_t)c->slot_cap * 8, c->stream));
    CUC(cudaMalloc(&g.blk_rec[L], (size_t)cfg->max_blocks * sizeof(BlockRec)));
    CUC(cudaMemsetAsync(g.blk_rec[L], 0, (size_t)cfg->max_blocks * sizeof(BlockRec), c->stream));
  }
  CUC(cudaMalloc(&c->spare_slots, (size_t)c->slot_cap * 8));
  CUC(cudaMalloc(&g.mdl_root, (size_t)cfg->max_models * 4));
  CUC(cudaMalloc(&g.mdl_ranger_return, (size_t)cfg->max_models * 8));
  CUC(cudaMalloc(&g.mdl_flags, (size_t)cfg->max_models * 4));
  CUC(cudaMalloc(&g.mdl_pal, (size_t)cfg->max_models));
  CUC(cudaMemsetAsync(g.mdl_pal, 0, (size_t)cfg->max_models, c->stream));
  CUC(cudaMemsetAsync(g.mdl_root, 0, (size_t)cfg->max_models * 4, c->stream));
  CUC(cudaMemsetAsync(g.mdl_ranger_return, 0, (size_t)cfg->max_models * 8, c->stream));
  CUC(cudaMemsetAsync(g.mdl_flags, 0, (size_t)cfg->max_models * 4, c->stream));
  CUC(cudaMalloc(&c->d_work_counter.p, 64));
  c->d_work_counter.cap = 64;
  CUC(cudaMallocHost(&c->err_host, 64));
  *c->err_host = 0;
  g.err_host = c->err_host;
  c->blob_cap = cfg->max_delta_bytes ? std::max<size_t>(cfg->max_delta_bytes, 4096) : (8u << 20);
  CUC(cudaMalloc(&c->d_blob.p, c->blob_cap));
  c->d_blob.cap = c->blob_cap;
  CUC(cudaMallocHost(&c->h_blob.p, c->blob_cap));
  c->h_blob.cap = c->blob_cap;
  CUC(cudaStreamSynchronize(c->stream));
#undef CUC
  c->blocks.resize(cfg->max_blocks);
  c->model_known.assign(cfg->max_models, false);
  c->models.resize(cfg->max_models);
  c->mdl_color.assign(4 * (size_t)cfg->max_models, 0.0);
  *out = c;
  return STG_RT_OK;
}

int stg_rt_destroy(stg_rt_ctx *c)
{
  if (!c)
    return STG_RT_EINVAL;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  if (c->prof_on)
    fprintf(stderr, "stg_rt host profile, total us (%llu fan launches): stage fans + headings %.0f | sort units %.0f | size deltas %.0f | "
                    "write deltas %.0f | upload + K1 launches %.0f | march launch %.0f | blocks_remap calls %.0f | fiducial unpack %.0f | "
                    "waiting for the GPU in sync %.0f | delivering fan results (copies, intensity expansion) %.0f\n",
            (unsigned long long)c->prof_flushes, c->prof_us[0], c->prof_us[1], c->prof_us[2], c->prof_us[3], c->prof_us[4],
            c->prof_us[5], c->prof_us[6], c->prof_us[7], c->prof_us[8], c->prof_us[9]);
  GridView &g = c->g;
  void *dev[] = { g.head, g.occ[0], g.slots[0], g.slots[1], g.blk_rec[0], g.blk_rec[1], c->spare_slots,
                  g.mdl_root, g.mdl_ranger_return, g.mdl_flags, g.mdl_pal, c->d_blob.p };
  for (void *p : dev)
    if (p)
      cudaFree(p);
  free_fan_arrays(c->d);
  free_fan_arrays(c->bd);
  if (c->q_fans.p)
    cudaFreeHost(c->q_fans.p);
  if (c->d_in.p)
    cudaFree(c->d_in.p);
  for (DevBuf *b : { &c->d_work_counter, &c->d_geom, &c->d_pts, &c->d_pix[0], &c->d_pix[1], &c->d_pix_new[0], &c->d_pix_new[1], &c->d_rig, &c->d_mv_in, &c->d_mv_out, &c->d_mover_of_model, &c->d_col_in, &c->d_col_hit, &c->d_touch_out, &c->d_fid_grid, &c->d_fid_in, &c->d_fid_out, &c->d_fid_count, &c->d_fid_off, &c->d_blob_in, &c->d_blob_out, &c->d_blob_count, &c->d_mdl_color, &c->d_stretch_count })
    if (b->p)
      cudaFree(b->p);
  for (PinBuf *b : { &c->h_rig, &c->h_mv_in, &c->h_mv_out, &c->h_col_in, &c->h_col_hit, &c->h_touch_out, &c->h_fid_in, &c->h_fid_out, &c->h_fid_count, &c->h_blob_in, &c->h_blob_out, &c->h_blob_count })
    if (b->p)
      cudaFreeHost(b->p);
  PinBuf *pins[] = { &c->h_blob, &c->h_in, &c->h_ranges, &c->h_intens, &c->h_bearings, &c->h_hit_model, &c->h_hit_cell, &c->h_steps,
                     &c->h_intens_idx, &c->h_done_flags };
  for (PinBuf *p : pins)
    if (p->p)
      cudaFreeHost(p->p);
  for (auto &kv : c->host_allocs)
    cudaFreeHost(kv.first);
  cudaEventDestroy(c->blob_uploaded);
  cudaEventDestroy(c->fans_uploaded);
  for (int k = 0; k < 8; k++)
    cudaEventDestroy(c->chunk_ev[k]);
  cudaEventDestroy(c->copies_done);
  cudaEventDestroy(c->hdrs_ready);
  cudaEventDestroy(c->xstream_ev);
  if (c->h_hdrs.p)
    cudaFreeHost(c->h_hdrs.p);
  if (c->err_host)
    cudaFreeHost(c->err_host);
  cudaStreamDestroy(c->copy_stream);
  if (c->sum_stream) {
    cudaStreamSynchronize(c->sum_stream);
    cudaStreamDestroy(c->sum_stream);
  }
  if (c->sum_ready)
    cudaEventDestroy(c->sum_ready);
  if (c->sum_done)
    cudaEventDestroy(c->sum_done);
  for (int k = 0; k < 6; k++)
    for (int j = 0; j < 2; j++)
      cudaEventDestroy(c->t_ev[k][j]);
  cudaStreamDestroy(c->own_stream);
  delete c;
  return STG_RT_OK;
}

int stg_rt_model_upsert(stg_rt_ctx *c, uint32_t id, uint32_t root_id, double ranger_return, int32_t fiducial_return,
                        int32_t fiducial_key, uint32_t vis_flags, const double *rgba)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (id < c->cfg.max_models && rgba) {
    memcpy(&c->mdl_color[4 * (size_t)id], rgba, 4 * sizeof(double));
    c->mdl_color_dirty = true;
  }
  if (id >= c->cfg.max_models || root_id >= c->cfg.max_models)
    return fail(c, STG_RT_ENOMEM, "model id %u / root %u beyond max_models %u", id, root_id, c->cfg.max_models);
  if (!c->q_subs.empty()) { // fans queued before this change must not see it
    int rc = flush_locked(c);
    if (rc)
      return rc;
  }
  ModelUpd u;
  memset(&u, 0, sizeof(u));
  u.id = id;
  u.root = root_id;
  u.ranger_return = ranger_return;
  u.fiducial_return = fiducial_return;
  u.fiducial_key = fiducial_key;
  u.flags = vis_flags;
  u.pal = 0;
  if (c->palette_ok) {
    size_t k = 1;
    while (k < c->palette.size() && memcmp(&c->palette[k], &ranger_return, sizeof(double)))
      k++;
    if (k == c->palette.size()) {
      if (k < kPaletteUnwritten) // (index 255 is never a palette entry: it marks a byte the march has not written yet)
        c->palette.push_back(ranger_return);
      else
        c->palette_ok = false; // more distinct values than a byte names: intensities travel as doubles from now on
    }
    u.pal = c->palette_ok ? (uint32_t)k : 0u;
  }
  c->model_upd.push_back(u);
  if (c->model_known[id] && (c->models[id].root != u.root || c->models[id].flags != u.flags ||
                             (c->models[id].ranger_return < 0) != (u.ranger_return < 0)))
    c->models_changed.push_back(id);
  c->models[id] = u;
  c->model_known[id] = true;
  if (sizeof(DeltaHeader) + c->model_upd.size() * sizeof(ModelUpd) > c->blob_cap / 2)
    return apply_deltas_locked(c);
  return STG_RT_OK;
}

static int block_map_locked(stg_rt_ctx *c, uint32_t block_id, uint32_t model_id, int layer, double zmin, double zmax, int n_pts,
                            const int32_t *xy);
static int block_unmap_locked(stg_rt_ctx *c, uint32_t block_id, int layer);

// The outline of (block, layer) as the device rendered it from a pose comes home: afterwards the host shadow holds it
// like an outline the caller supplied (needed by callers that re-derive a block's cells from its outline: collision
// tests, touching models, batched moves, a second Map on top). Applies what is queued first.
static int fetch_outline_locked(stg_rt_ctx *c, uint32_t block_id, int layer)
{
  int rc = flush_locked(c);
  if (rc)
    return rc;
  BlockShadow &b = c->blocks[block_id];
  if (!b.dev_mapped[layer] || !b.dev_owned[layer])
    return STG_RT_OK;
  const GeomShadow &g = c->geom[block_id];
  b.vec_used[layer] = true;
  b.dev_xy[layer].resize(2 * (size_t)g.n_pts);
  CU(c, cudaMemcpyAsync(b.dev_xy[layer].data(), (const int32_t *)c->d_pix[layer].p + 2 * (size_t)g.first_pt, (size_t)g.n_pts * 8,
                        cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->stats.d2h_bytes += (size_t)g.n_pts * 8;
  uint64_t steps = 0;
  const std::vector<int32_t> &xy = b.dev_xy[layer];
  for (uint32_t i = 0; i < g.n_pts; i++) {
    const uint32_t j = i + 1 == g.n_pts ? 0 : i + 1;
    steps += (uint64_t)std::llabs((long long)xy[2 * j] - xy[2 * i]) + (uint64_t)std::llabs((long long)xy[2 * j + 1] - xy[2 * i + 1]);
  }
  c->live_entries[layer] += (int64_t)steps - (int64_t)b.dev_steps[layer]; // the estimate gives way to the count
  b.dev_steps[layer] = (uint32_t)steps;
  b.dev_poly[layer].clear();
  b.dev_owned[layer] = false;
  if (b.host_mapped[layer] && !b.dirty[layer])
    b.host_steps[layer] = b.dev_steps[layer];
  return STG_RT_OK;
}

int stg_rt_block_map(stg_rt_ctx *c, uint32_t block_id, uint32_t model_id, int layer, double zmin, double zmax, int n_pts,
                     const int32_t *xy)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  return block_map_locked(c, block_id, model_id, layer, zmin, zmax, n_pts, xy);
}

int stg_rt_block_unmap(stg_rt_ctx *c, uint32_t block_id, int layer)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  return block_unmap_locked(c, block_id, layer);
}

// World::AddSuperRegion (world.cc:1072-1093): the reference's world grows whenever something is rendered outside it. Here
// the dense extent is re-laid over the larger rectangle of SuperRegions that holds the old one and the cells
// [x0..x1] x [y0..y1] (plus one SuperRegion of room on the sides that grew), on the device (launch_grow).
static int grow_extent_locked(stg_rt_ctx *c, int32_t x0, int32_t y0, int32_t x1, int32_t y1)
{
  GridView &g = c->g;
  const int32_t osx0 = g.rx0 >> 5, osy0 = g.ry0 >> 5, osx1 = osx0 + (g.nrx >> 5) - 1, osy1 = osy0 + (g.nry >> 5) - 1;
  int32_t sx0 = std::min(osx0, x0 >> 10), sy0 = std::min(osy0, y0 >> 10), sx1 = std::max(osx1, x1 >> 10), sy1 = std::max(osy1, y1 >> 10);
  if (sx0 < osx0) sx0--; // room to keep going that way
  if (sy0 < osy0) sy0--;
  if (sx1 > osx1) sx1++;
  if (sy1 > osy1) sy1++;
  const int64_t nsx = (int64_t)sx1 - sx0 + 1, nsy = (int64_t)sy1 - sy0 + 1;
  if (nsx * nsy > 4095)
    return fail(c, STG_RT_EEXTENT, "the world would span %lld x %lld SuperRegions; the device grid holds at most 4095", (long long)nsx, (long long)nsy);
  int rc = flush_locked(c); // what is queued belongs to the grid as it is
  if (rc || (rc = deliver_locked(c)))
    return rc;
  GridView to = g;
  to.rx0 = sx0 * 32;
  to.ry0 = sy0 * 32;
  to.nrx = (int32_t)nsx * 32;
  to.nry = (int32_t)nsy * 32;
  const size_t n_new = (size_t)to.nrx * to.nry;
  to.n_regions = (uint32_t)n_new;
  to.head = nullptr;
  to.occ[0] = nullptr;
  to.slots[0] = to.slots[1] = nullptr;
  unsigned long long *fresh[2] = { nullptr, nullptr };
  cudaError_t e = cudaMalloc(&to.head, (n_new + 1) * sizeof(RegionHead));
  if (e == cudaSuccess) e = cudaMalloc(&to.occ[0], occ_alloc_words(n_new) * 4);
  for (int L = 0; L < 2 && e == cudaSuccess; L++)
    e = cudaMalloc(&fresh[L], (size_t)c->slot_cap * 8);
  if (e == cudaSuccess) e = cudaMemsetAsync(to.head, 0, (n_new + 1) * sizeof(RegionHead), c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(to.occ[0], 0, occ_alloc_words(n_new) * 4, c->stream);
  for (int L = 0; L < 2 && e == cudaSuccess; L++)
    e = cudaMemsetAsync(fresh[L], 0, (size_t)c->slot_cap * 8, c->stream);
  if (e != cudaSuccess) {
    for (void *p : { (void *)to.head, (void *)to.occ[0], (void *)fresh[0], (void *)fresh[1] })
      if (p)
        cudaFree(p);
    cudaGetLastError();
    return fail(c, STG_RT_ENOMEM, "no device memory for a grid of %lld x %lld SuperRegions: %s", (long long)nsx, (long long)nsy, cudaGetErrorString(e));
  }
  to.occ[1] = to.occ[0] + n_new * kTileWords;
  to.occ_dirty = to.occ[0] + 2 * n_new * kTileWords;
  to.slots[0] = fresh[0];
  to.slots[1] = fresh[1];
  CU(c, cudaStreamSynchronize(c->sum_stream)); // the old arrays are about to be read whole and freed
  launch_grow(g, to, c->stream);
  c->stats.launches_other += 3;
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  for (void *p : { (void *)g.head, (void *)g.occ[0], (void *)g.slots[0], (void *)g.slots[1] })
    cudaFree(p);
  g = to;
  c->n_regions = n_new;
  c->cell_x0 = g.rx0 * 32;
  c->cell_y0 = g.ry0 * 32;
  c->cell_x1 = c->cell_x0 + g.nrx * 32 - 1;
  c->cell_y1 = c->cell_y0 + g.nry * 32 - 1;
  c->cfg.sreg_x0 = sx0;
  c->cfg.sreg_y0 = sy0;
  c->cfg.sreg_nx = (int32_t)nsx;
  c->cfg.sreg_ny = (int32_t)nsy;
  for (int L = 0; L < 2; L++)
    c->used_upper[L] = c->live_entries[L]; // fresh tables: no tombstones
  c->stats.extent_grows++;
  return STG_RT_OK;
}

int stg_rt_block_define(stg_rt_ctx *c, uint32_t block_id, uint32_t model_id, double local_zmin, double local_zmax, int n_pts,
                        const double *pts_xy)
{
  if (!c || n_pts < 1 || !pts_xy)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (block_id >= c->cfg.max_blocks)
    return fail(c, STG_RT_ENOMEM, "block id %u beyond max_blocks %u", block_id, c->cfg.max_blocks);
  if (model_id >= c->cfg.max_models || !c->model_known[model_id])
    return fail(c, STG_RT_EINVAL, "block_define: unknown model %u (call stg_rt_model_upsert first)", model_id);
  BlockShadow &b = c->blocks[block_id];
  for (int L = 0; L < 2; L++)
    if ((b.dev_mapped[L] && b.dev_owned[L]) || (b.host_mapped[L] && b.pose_pending[L]))
      return fail(c, STG_RT_ESTATE, "block_define: block %u is rendered from its present definition on layer %d; unmap it first", block_id, L);
  if (c->geom.size() <= block_id)
    c->geom.resize((size_t)block_id + 1);
  GeomShadow &g = c->geom[block_id];
  if (g.defined && g.model != model_id) { // the block id now names a block of another model
    std::vector<uint32_t> &lst = c->model_blocks[g.model];
    lst.erase(std::remove(lst.begin(), lst.end(), block_id), lst.end());
    g.defined = false;
  }
  if (!g.defined || g.n_pts != (uint32_t)n_pts) { // a fresh stretch of the point pools
    g.first_pt = (uint32_t)(c->h_pts.size() / 2);
    c->h_pts.resize(c->h_pts.size() + 2 * (size_t)n_pts);
  }
  if (!g.defined) {
    if (c->model_blocks.size() <= model_id)
      c->model_blocks.resize((size_t)model_id + 1);
    c->model_blocks[model_id].push_back(block_id);
  }
  g.defined = true;
  g.n_pts = (uint32_t)n_pts;
  g.model = model_id;
  g.zmin = local_zmin;
  g.zmax = local_zmax;
  memcpy(&c->h_pts[2 * (size_t)g.first_pt], pts_xy, 2 * (size_t)n_pts * sizeof(double));
  double r2 = 0, perimeter = 0;
  for (int i = 0; i < n_pts; i++) {
    const int j = i + 1 == n_pts ? 0 : i + 1;
    r2 = std::max(r2, pts_xy[2 * i] * pts_xy[2 * i] + pts_xy[2 * i + 1] * pts_xy[2 * i + 1]);
    perimeter += fabs(pts_xy[2 * j] - pts_xy[2 * i]) + fabs(pts_xy[2 * j + 1] - pts_xy[2 * i + 1]);
  }
  g.radius = sqrt(r2);
  // |dx| + |dy| of a rotated edge is at most sqrt(2) times its length, and floor() moves each end by less than a cell
  g.est_steps = (uint32_t)std::min(1.0e9, ceil(1.4142135623730951 * perimeter * c->cfg.ppm) + 2.0 * n_pts);
  b.known = true;
  if (b.model != model_id && (b.host_mapped[0] || b.host_mapped[1] || b.dev_mapped[0] || b.dev_mapped[1]))
    c->block_reassigned = true;
  b.model = model_id;
  c->geom_dirty = true;
  return STG_RT_OK;
}

int stg_rt_models_pose_update(stg_rt_ctx *c, uint32_t n, const uint32_t *models, const double *gpose_xyza, int layer)
{
  if (!c || layer < 0 || layer > 1 || (n && (!models || !gpose_xyza)))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  const double t_prof = prof_now(c);
  // A swarm's tick: many models of one block each, in ascending id order, well inside the extent, none of them queued yet,
  // no fans waiting for the grid as it was. The host threads first make sure of all that side by side, then queue the
  // poses side by side (model i's Map gets sequence number local_seq + i + 1 and unit number queue length + i, as the
  // loop below would give them). Anything else takes the loop.
  if (n >= 8192 && c->q_subs.empty() && (uint64_t)c->local_seq + n < kSeqLocalMax && !c->pool.busy()) {
    std::atomic<bool> plain{ true };
    const double ppm = c->cfg.ppm;
    c->pool.parallel_for(n, 4096, [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++) {
        const uint32_t m = models[i];
        if ((i && models[i - 1] >= m) || m >= c->model_blocks.size() || c->model_blocks[m].size() != 1) {
          plain.store(false, std::memory_order_relaxed);
          return;
        }
        const uint32_t block_id = c->model_blocks[m][0];
        const GeomShadow &g = c->geom[block_id];
        const double *P = gpose_xyza + 4 * i, reach = (g.radius + 2.0 / ppm) * ppm;
        if (!(P[0] * ppm - reach >= c->cell_x0 && P[0] * ppm + reach <= c->cell_x1 + 1 && P[1] * ppm - reach >= c->cell_y0 &&
              P[1] * ppm + reach <= c->cell_y1 + 1) || c->blocks[block_id].dirty[layer]) {
          plain.store(false, std::memory_order_relaxed);
          return;
        }
      }
    });
    if (plain.load()) {
      const uint32_t seq0 = c->local_seq;
      const size_t unit0 = c->dirty_units.size();
      if (unit0 == 0)
        c->units_in_order = true;
      else if (seq0 + 1 < c->last_unit_key)
        c->units_in_order = false;
      c->dirty_units.resize(unit0 + n);
      uint32_t *units = c->dirty_units.data() + unit0;
      c->pool.parallel_for(n, 4096, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
          const uint32_t block_id = c->model_blocks[models[i]][0];
          const GeomShadow &g = c->geom[block_id];
          const double *P = gpose_xyza + 4 * i;
          BlockShadow &b = c->blocks[block_id];
          b.zmin = g.zmin + P[2];
          b.zmax = g.zmax + P[2];
          b.host_mapped[layer] = true;
          b.pose_pending[layer] = true;
          memcpy(b.pose[layer], P, 4 * sizeof(double));
          b.host_steps[layer] = g.est_steps;
          if (b.vec_used[layer]) {
            b.host_xy[layer].clear();
            b.host_poly[layer].clear();
          }
          b.keep_seq[layer] = false;
          b.map_seq[layer] = seq0 + (uint32_t)i + 1u;
          b.dirty[layer] = true;
          units[i] = block_id * 2 + (uint32_t)layer;
        }
      });
      c->local_seq = seq0 + n;
      c->last_unit_key = seq0 + n;
      c->stats.map_calls += n;
      c->stats.unmap_calls += n;
      if (c->prof_on)
        c->prof_us[6] += prof_now(c) - t_prof;
      return STG_RT_OK;
    }
  }
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t m = models[i];
    if (m >= c->model_blocks.size() || c->model_blocks[m].empty())
      return fail(c, STG_RT_EINVAL, "models_pose_update: model %u has no defined blocks (stg_rt_block_define)", m);
    const double *P = gpose_xyza + 4 * (size_t)i;
    const std::vector<uint32_t> &lst = c->model_blocks[m];
    for (size_t k = 0; k < lst.size(); k++) { // BlockGroup::UnMap then BlockGroup::Map: per block, in BlockGroup order
      const uint32_t block_id = lst[k];
      const GeomShadow &g = c->geom[block_id];
      // every vertex lies within `radius` of the model origin: a pose this far inside the extent cannot leave it
      const double reach = (g.radius + 2.0 / c->cfg.ppm) * c->cfg.ppm;
      if (!(P[0] * c->cfg.ppm - reach >= c->cell_x0 && P[0] * c->cfg.ppm + reach <= c->cell_x1 + 1 && P[1] * c->cfg.ppm - reach >= c->cell_y0 &&
            P[1] * c->cfg.ppm + reach <= c->cell_y1 + 1)) {
        // too close to the edge for the bound to decide (a floorplan as large as the world): the vertices themselves, as
        // Model::LocalToPixels places them
        const double cosa = cos(P[3]), sina = sin(P[3]);
        const double *pt = &c->h_pts[2 * (size_t)g.first_pt];
        for (uint32_t v = 0; v < g.n_pts; v++) {
          const double px = floor((P[0] + pt[2 * v] * cosa - pt[2 * v + 1] * sina) * c->cfg.ppm);
          const double py = floor((P[1] + pt[2 * v] * sina + pt[2 * v + 1] * cosa) * c->cfg.ppm);
          if (!(px >= c->cell_x0 && px <= c->cell_x1 && py >= c->cell_y0 && py <= c->cell_y1)) {
            if (!(fabs(px) < 2.0e9 && fabs(py) < 2.0e9))
              return fail(c, STG_RT_EEXTENT, "model %u at (%.3f, %.3f): block %u vertex beyond 32-bit cell coordinates", m, P[0], P[1], block_id);
            const int rcg = grow_extent_locked(c, (int32_t)px, (int32_t)py, (int32_t)px, (int32_t)py); // the world grows
            if (rcg)
              return rcg;
          }
        }
      }
      if (!c->q_subs.empty()) { // issue order: queued fans must be traced against the grid as it was
        int rc = flush_locked(c);
        if (rc)
          return rc;
      }
      if (c->local_seq >= kSeqLocalMax) {
        if (c->export_mode)
          return fail(c, STG_RT_ENOMEM, "more than %u Map calls between two delta exchanges", kSeqLocalMax);
        int rc = apply_deltas_locked(c);
        if (rc)
          return rc;
      }
      BlockShadow &b = c->blocks[block_id];
      b.zmin = g.zmin + P[2]; // Block::Map, block.cc:177-180 (what the host-side collision queries read)
      b.zmax = g.zmax + P[2];
      b.host_mapped[layer] = true;
      b.pose_pending[layer] = true;
      memcpy(b.pose[layer], P, 4 * sizeof(double));
      b.host_steps[layer] = g.est_steps;
      if (b.vec_used[layer]) {
        b.host_xy[layer].clear();
        b.host_poly[layer].clear();
      }
      b.keep_seq[layer] = false;
      b.map_seq[layer] = ++c->local_seq;
      touch_unit(c, block_id, layer);
      c->stats.map_calls++;
      c->stats.unmap_calls++;
    }
  }
  if (c->prof_on)
    c->prof_us[6] += prof_now(c) - t_prof;
  return STG_RT_OK;
}

int stg_rt_debug_block_outline(stg_rt_ctx *c, uint32_t block_id, int layer, int32_t *xy, int cap_pts)
{
  if (!c || layer < 0 || layer > 1 || block_id >= c->cfg.max_blocks)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc || (rc = deliver_locked(c)))
    return rc;
  const BlockShadow &b = c->blocks[block_id];
  if (!b.dev_mapped[layer])
    return 0;
  if (!b.dev_owned[layer]) {
    const int n = (int)(b.dev_xy[layer].size() / 2);
    if (xy)
      memcpy(xy, b.dev_xy[layer].data(), (size_t)std::min(n, cap_pts) * 8);
    return n;
  }
  const GeomShadow &g = c->geom[block_id];
  if (xy && cap_pts > 0) {
    CU(c, cudaMemcpyAsync(xy, (const int32_t *)c->d_pix[layer].p + 2 * (size_t)g.first_pt, (size_t)std::min<uint32_t>(g.n_pts, (uint32_t)cap_pts) * 8,
                          cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return (int)g.n_pts;
}

int stg_rt_blocks_remap(stg_rt_ctx *c, uint32_t n, const uint32_t *blocks, const uint32_t *models, int layer, const double *z,
                        const uint32_t *first_pt, const int32_t *xy)
{
  if (!c || (n && (!blocks || !models || !z || !first_pt || !xy)))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  const double t_prof = prof_now(c);
  for (uint32_t i = 0; i < n; i++) {
    int rc = block_unmap_locked(c, blocks[i], layer);
    if (!rc)
      rc = block_map_locked(c, blocks[i], models[i], layer, z[2 * i], z[2 * i + 1], (int)(first_pt[i + 1] - first_pt[i]),
                            xy + 2 * (size_t)first_pt[i]);
    if (rc)
      return rc;
  }
  if (c->prof_on)
    c->prof_us[6] += prof_now(c) - t_prof;
  return STG_RT_OK;
}

static int block_map_locked(stg_rt_ctx *c, uint32_t block_id, uint32_t model_id, int layer, double zmin, double zmax, int n_pts,
                            const int32_t *xy)
{
  if (layer < 0 || layer > 1 || n_pts < 0 || (n_pts && !xy))
    return fail(c, STG_RT_EINVAL, "block_map: bad layer / points");
  if (block_id >= c->cfg.max_blocks)
    return fail(c, STG_RT_ENOMEM, "block id %u beyond max_blocks %u", block_id, c->cfg.max_blocks);
  if (model_id >= c->cfg.max_models || !c->model_known[model_id])
    return fail(c, STG_RT_EINVAL, "block_map: unknown model %u (call stg_rt_model_upsert first)", model_id);
  BlockShadow &b = c->blocks[block_id];
  uint64_t steps = 0; // cell visits of this outline: |dx| + |dy| per edge (world.cc:1000-1002)
  {
    int32_t bx0 = c->cell_x0, by0 = c->cell_y0, bx1 = c->cell_x1, by1 = c->cell_y1;
    for (int i = 0; i < n_pts; i++) {
      bx0 = std::min(bx0, xy[2 * i]);
      bx1 = std::max(bx1, xy[2 * i]);
      by0 = std::min(by0, xy[2 * i + 1]);
      by1 = std::max(by1, xy[2 * i + 1]);
    }
    if (bx0 < c->cell_x0 || by0 < c->cell_y0 || bx1 > c->cell_x1 || by1 > c->cell_y1) { // the world grows, as the reference's does
      const int rcg = grow_extent_locked(c, bx0, by0, bx1, by1);
      if (rcg)
        return rcg;
    }
  }
  for (int i = 0; i < n_pts; i++) {
    const int j = i + 1 == n_pts ? 0 : i + 1;
    steps += (uint64_t)std::llabs((long long)xy[2 * j] - xy[2 * i]) + (uint64_t)std::llabs((long long)xy[2 * j + 1] - xy[2 * i + 1]);
  }
  if (steps + b.host_steps[layer] > 0x7fffffffull)
    return fail(c, STG_RT_EINVAL, "block %u: outline of %llu cell visits", block_id, (unsigned long long)steps);
  if (!c->q_subs.empty()) { // issue order: queued fans must be traced against the grid as it was
    int rc = flush_locked(c);
    if (rc)
      return rc;
  }
  if (c->local_seq >= kSeqLocalMax) { // the call-order field of a sequence number is full: close this round of deltas
    if (c->export_mode)
      return fail(c, STG_RT_ENOMEM, "more than %u Map calls between two delta exchanges", kSeqLocalMax);
    int rc = apply_deltas_locked(c);
    if (rc)
      return rc;
  }
  if (b.host_mapped[layer] && ((b.dirty[layer] && b.pose_pending[layer]) || (!b.dirty[layer] && b.dev_owned[layer]))) {
    // one more rendering on top of an outline only the device knows: bring that outline home first
    int rc = fetch_outline_locked(c, block_id, layer);
    if (rc)
      return rc;
  }
  b.pose_pending[layer] = false;
  b.vec_used[layer] = true;
  if (b.known && b.model != model_id)
    c->block_reassigned = true; // what is still rendered of it on the other layer changes owner too
  b.known = true;
  b.model = model_id;
  b.zmin = zmin;
  b.zmax = zmax;
  if (b.host_mapped[layer]) {
    // Map on a mapped block (see BlockShadow): one more rendering on top of what is there; the block keeps the
    // place in the cell lists its first rendering gave it
    if (!b.dirty[layer]) { // the device holds the current outline; host_xy is scratch since the last flush
      b.host_xy[layer] = b.dev_xy[layer];
      b.host_poly[layer] = b.dev_poly[layer];
      b.host_steps[layer] = b.dev_steps[layer];
      b.keep_seq[layer] = true;
    }
    if (b.host_poly[layer].empty())
      b.host_poly[layer].push_back((uint32_t)(b.host_xy[layer].size() / 2));
    b.host_poly[layer].push_back((uint32_t)n_pts);
    b.host_xy[layer].insert(b.host_xy[layer].end(), xy, xy + 2 * (size_t)n_pts);
    b.host_steps[layer] += (uint32_t)steps;
    c->local_seq++;
  } else {
    b.host_mapped[layer] = true;
    b.host_steps[layer] = (uint32_t)steps;
    b.host_xy[layer].assign(xy, xy + 2 * (size_t)n_pts);
    b.host_poly[layer].clear();
    b.keep_seq[layer] = false;
    b.map_seq[layer] = ++c->local_seq;
  }
  touch_unit(c, block_id, layer);
  c->stats.map_calls++;
  return STG_RT_OK;
}

static int block_unmap_locked(stg_rt_ctx *c, uint32_t block_id, int layer)
{
  if (layer < 0 || layer > 1)
    return fail(c, STG_RT_EINVAL, "block_unmap: bad layer");
  if (block_id >= c->cfg.max_blocks)
    return fail(c, STG_RT_ENOMEM, "block id %u beyond max_blocks %u", block_id, c->cfg.max_blocks);
  BlockShadow &b = c->blocks[block_id];
  if (!b.host_mapped[layer])
    return STG_RT_OK;
  if (!c->q_subs.empty()) {
    int rc = flush_locked(c);
    if (rc)
      return rc;
  }
  b.host_mapped[layer] = false;
  b.pose_pending[layer] = false;
  b.host_steps[layer] = 0;
  b.host_xy[layer].clear();
  b.host_poly[layer].clear();
  b.keep_seq[layer] = false;
  touch_unit(c, block_id, layer);
  c->stats.unmap_calls++;
  return STG_RT_OK;
}

static int fans_submit_locked(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const double *headings, const double *trig,
                              const stg_rt_fan_noise *noise, const stg_rt_fan_out *out);

int stg_rt_fans_submit(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const double *headings, const double *trig,
                       const stg_rt_fan_out *out)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  return fans_submit_locked(c, n_fans, fans, headings, trig, nullptr, out);
}

int stg_rt_fans_submit_noisy(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const stg_rt_fan_noise *noise,
                             const stg_rt_fan_out *out)
{
  static_assert(sizeof(stg_rt_fan_noise) == sizeof(FanNoiseRec), "noise record layout");
  if (!c || (n_fans && !noise))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  for (uint32_t i = 0; i < n_fans; i++)
    if (fans && fans[i].angles != STG_RT_ANGLES_RANGER)
      return fail(c, STG_RT_EINVAL, "fans_submit_noisy: fan %u is not a ranger fan", i);
  return fans_submit_locked(c, n_fans, fans, nullptr, nullptr, noise, out);
}

static int fans_submit_locked(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const double *headings, const double *trig,
                              const stg_rt_fan_noise *noise, const stg_rt_fan_out *out)
{
  if (!n_fans)
    return STG_RT_OK;
  if (!fans || !out)
    return fail(c, STG_RT_EINVAL, "fans_submit: null fans / out");
  int rc = 0;
  if (c->q_rig.active && (rc = flush_locked(c))) // a ranger-model batch is a batch of its own
    return rc;
  if (!c->q_subs.empty() && c->q_has_trig != (trig != nullptr)) { // a batch is all-strict or all-device trig
    if ((rc = flush_locked(c)))
      return rc;
  }
  Submit s;
  s.first_beam = c->q_beams;
  s.out = *out;
  uint64_t beams = 0;
  const size_t head_base = c->q_headings.size(), fan_base = c->q_fans.size();
  uint32_t max_head = 0;
  static_assert(sizeof(stg_rt_fan) == sizeof(FanRec), "fan record layout");
  if (!fan_base)
    CU(c, cudaEventSynchronize(c->fans_uploaded)); // the previous batch has left the queue's memory
  if ((rc = pinfans_resize(c, fan_base + n_fans)))
    return rc;
  c->q_first.resize(fan_base + n_fans);
  FanRec *qf = c->q_fans.data() + fan_base;
  uint32_t *first = c->q_first.data() + fan_base;
  // one pass over the records (a batch of sonars is a million of them): check, copy, book-keep
  for (uint32_t i = 0; i < n_fans && !rc; i++) {
    const stg_rt_fan &f = fans[i];
    if (f.finder >= c->cfg.max_models || !c->model_known[f.finder])
      rc = fail(c, STG_RT_EINVAL, "fan %u: unknown finder model %u", i, f.finder);
    else if (f.pred > STG_RT_PRED_GRIPPER || f.angles > STG_RT_ANGLES_EXPLICIT || f.layer > 1)
      rc = fail(c, STG_RT_EINVAL, "fan %u: bad pred/angles/layer", i);
    else if (f.angles == STG_RT_ANGLES_EXPLICIT && !headings)
      rc = fail(c, STG_RT_EINVAL, "fan %u: explicit headings missing", i);
    memcpy(&qf[i], &f, sizeof(FanRec));
    first[i] = (uint32_t)(c->q_beams + beams);
    if (f.angles == STG_RT_ANGLES_EXPLICIT) {
      max_head = std::max(max_head, f.first_heading + f.n_samples);
      qf[i].first_heading += (uint32_t)head_base;
    }
    beams += f.n_samples;
  }
  if (!rc && c->q_beams + beams >= (1ull << 32))
    rc = fail(c, STG_RT_EINVAL, "more than 2^32 beams in one batch");
  if (rc) {
    c->q_fans.n = fan_base;
    c->q_first.resize(fan_base);
    return rc;
  }
  if (max_head)
    c->q_headings.insert(c->q_headings.end(), headings, headings + max_head);
  if (trig) {
    c->q_trig.insert(c->q_trig.end(), trig, trig + 2 * beams);
    c->q_has_trig = true;
  }
  if (noise) {
    FanNoiseRec zero;
    memset(&zero, 0, sizeof(zero));
    c->q_noise.resize(c->q_fans.size() - n_fans, zero); // earlier fans of this batch carry none
    for (uint32_t i = 0; i < n_fans; i++) {
      FanNoiseRec r;
      memcpy(&r, &noise[i], sizeof(r));
      c->q_noise.push_back(r);
    }
    c->q_has_noise = true;
  }
  s.n_beams = beams;
  c->q_beams += beams;
  c->q_subs.push_back(s);
  return STG_RT_OK;
}

int stg_rt_rangers_submit(stg_rt_ctx *c, uint32_t n_sensors, const stg_rt_ranger_sensor *sensors, uint32_t n_rangers,
                          const stg_rt_ranger_pose *poses, const stg_rt_fan_out *out)
{
  static_assert(sizeof(stg_rt_ranger_sensor) == sizeof(RangerSensorRec) && sizeof(stg_rt_ranger_pose) == sizeof(RangerPoseRec),
                "ranger record layouts");
  if (!c || !out || (n_sensors && !sensors) || (n_rangers && !poses))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!n_sensors || !n_rangers)
    return STG_RT_OK;
  uint64_t per_model = 0;
  for (uint32_t s = 0; s < n_sensors; s++)
    per_model += sensors[s].n_samples;
  if ((uint64_t)n_rangers * n_sensors >= (1ull << 32) || per_model * n_rangers >= (1ull << 32))
    return fail(c, STG_RT_EINVAL, "rangers_submit: more than 2^32 fans or beams");
  for (uint32_t r = 0; r < n_rangers; r++)
    if (poses[r].finder >= c->cfg.max_models || !c->model_known[poses[r].finder] || poses[r].layer > 1)
      return fail(c, STG_RT_EINVAL, "ranger %u: unknown model %u or bad layer", r, poses[r].finder);
  int rc = flush_locked(c); // a batch of its own
  if (rc || (rc = deliver_locked(c)))
    return rc;
  stg_rt_ctx::RigJob &q = c->q_rig;
  q.n_sensors = n_sensors;
  q.n_rangers = n_rangers;
  q.o_first = (size_t)n_sensors * sizeof(RangerSensorRec);
  q.o_poses = (q.o_first + (size_t)(n_sensors + 1) * 4 + 15) & ~(size_t)15;
  q.bytes = q.o_poses + (size_t)n_rangers * sizeof(RangerPoseRec);
  if ((rc = pin_reserve(c, c->h_rig, q.bytes)))
    return rc;
  char *hp = (char *)c->h_rig.p;
  memcpy(hp, sensors, q.o_first);
  uint32_t *sf = (uint32_t *)(hp + q.o_first);
  q.uniform_samples = sensors[0].n_samples;
  uint32_t run = 0;
  for (uint32_t s = 0; s < n_sensors; s++) {
    sf[s] = run;
    run += sensors[s].n_samples;
    if (sensors[s].n_samples != sensors[0].n_samples)
      q.uniform_samples = 0;
  }
  sf[n_sensors] = run;
  memcpy(hp + q.o_poses, poses, (size_t)n_rangers * sizeof(RangerPoseRec));
  q.active = true;
  Submit s;
  s.first_beam = 0;
  s.n_beams = per_model * n_rangers;
  s.out = *out;
  c->q_beams = s.n_beams;
  c->q_subs.push_back(s);
  return STG_RT_OK;
}

int stg_rt_flush(stg_rt_ctx *c)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  return flush_locked(c);
}

int stg_rt_sync(stg_rt_ctx *c)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc)
    return rc;
  if ((rc = deliver_locked(c)))
    return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  return device_errors_locked(c);
}

void *stg_rt_host_alloc(stg_rt_ctx *c, size_t bytes)
{
  if (!c || !bytes)
    return nullptr;
  std::lock_guard<std::mutex> lk(c->mu);
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    fail(c, STG_RT_ECUDA, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  c->host_allocs[(char *)p] = bytes;
  return p;
}

int stg_rt_host_free(stg_rt_ctx *c, void *p)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->host_allocs.find((char *)p);
  if (it == c->host_allocs.end())
    return fail(c, STG_RT_EINVAL, "host_free: not a stg_rt_host_alloc pointer");
  deliver_locked(c);
  cudaFreeHost(p);
  c->host_allocs.erase(it);
  return STG_RT_OK;
}

// ---- resident sets ------------------------------------------------------------------------------

int stg_rt_set_create(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const double *headings, const double *trig,
                      uint32_t want, stg_rt_set **out)
{
  if (!c || !out || !fans || !n_fans)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = validate_fans(c, n_fans, fans, headings);
  if (rc)
    return rc;
  stg_rt_set *s = new stg_rt_set;
  s->want = want;
  uint64_t beams = 0;
  uint32_t max_head = 0;
  for (uint32_t i = 0; i < n_fans; i++) {
    FanRec r;
    memcpy(&r, &fans[i], sizeof(r));
    s->first.push_back((uint32_t)beams);
    if (r.angles == STG_RT_ANGLES_EXPLICIT)
      max_head = std::max(max_head, r.first_heading + r.n_samples);
    s->fans.push_back(r);
    beams += r.n_samples;
  }
  s->first.push_back((uint32_t)beams);
  s->n_beams = beams;
  s->has_trig = trig != nullptr;
  s->has_headings = max_head != 0;
  if ((rc = dev_reserve(c, s->d.fans, (size_t)n_fans * sizeof(FanRec))) || (rc = dev_reserve(c, s->d.first, (size_t)(n_fans + 1) * 4)) ||
      (rc = dev_reserve(c, s->d.headings_in, std::max<size_t>((size_t)max_head * 8, 8))) ||
      (rc = dev_reserve(c, s->d.trig, std::max<size_t>(trig ? beams * 16 : 0, 8))) || (rc = reserve_outputs(c, s->d, beams, want))) {
    free_fan_arrays(s->d);
    delete s;
    return rc;
  }
  CU(c, cudaMemcpyAsync(s->d.first.p, s->first.data(), (size_t)(n_fans + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  if (max_head)
    CU(c, cudaMemcpyAsync(s->d.headings_in.p, headings, (size_t)max_head * 8, cudaMemcpyHostToDevice, c->stream));
  if (trig)
    CU(c, cudaMemcpyAsync(s->d.trig.p, trig, beams * 16, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(s->d.fans.p, s->fans.data(), (size_t)n_fans * sizeof(FanRec), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream)); // sources are pageable caller memory
  s->fans_dirty = false;
  *out = s;
  return STG_RT_OK;
}

int stg_rt_set_destroy(stg_rt_ctx *c, stg_rt_set *s)
{
  if (!c || !s)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStreamSynchronize(c->stream);
  free_fan_arrays(s->d);
  delete s;
  return STG_RT_OK;
}

int stg_rt_set_update_origins(stg_rt_ctx *c, stg_rt_set *s, const double *xyza, int layer)
{
  if (!c || !s || !xyza || layer < 0 || layer > 1)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  for (size_t i = 0; i < s->fans.size(); i++) {
    FanRec &r = s->fans[i];
    r.ox = xyza[4 * i];
    r.oy = xyza[4 * i + 1];
    r.oz = xyza[4 * i + 2];
    r.oa = xyza[4 * i + 3];
    r.layer = (uint8_t)layer;
  }
  s->fans_dirty = true;
  return STG_RT_OK;
}

int stg_rt_set_trace(stg_rt_ctx *c, stg_rt_set *s)
{
  if (!c || !s)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc)
    return rc;
  if (s->fans_dirty) {
    CU(c, cudaStreamSynchronize(c->stream)); // the previous upload may still read s->fans
    CU(c, cudaMemcpyAsync(s->d.fans.p, s->fans.data(), s->fans.size() * sizeof(FanRec), cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += s->fans.size() * sizeof(FanRec);
    s->fans_dirty = false;
  }
  FanBatchView v = view_of(s->d, (uint32_t)s->fans.size(), s->n_beams, s->want, s->has_trig, s->has_headings);
  v.uniform_samples = uniform_samples_of(s->fans.data(), s->fans.size());
  const bool refill = ray_march_refill_wanted(v);
  if (refill && !v.trig_in && (rc = dev_reserve(c, s->d.trigw, s->n_beams * 16)))
    return rc;
  CU(c, cudaEventRecord(c->t_ev[1][0], c->stream));
  launch_fan_headings(v, c->stream);
  if (refill && !v.trig_in) {
    launch_beam_trig(v, (double *)s->d.trigw.p, c->stream);
    c->stats.launches_post++;
  }
  CU(c, cudaEventRecord(c->t_ev[1][1], c->stream));
  CU(c, cudaEventRecord(c->t_ev[2][0], c->stream));
  if (refill)
    launch_ray_march_refill(c->g, v, v.trig_in ? v.trig_in : (const double *)s->d.trigw.p, (uint32_t *)c->d_work_counter.p, c->stream);
  else
    launch_ray_march(c->g, v, c->stream);
  CU(c, cudaEventRecord(c->t_ev[2][1], c->stream));
  c->t_rec[1] = c->t_rec[2] = true;
  CU(c, cudaGetLastError());
  c->stats.launches_post += 1;
  c->stats.launches_march += 1;
  c->stats.beams += s->n_beams;
  c->stats.fans += s->fans.size();
  return STG_RT_OK;
}

int stg_rt_set_download(stg_rt_ctx *c, stg_rt_set *s, const stg_rt_fan_out *out)
{
  if (!c || !s || !out)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (want_of(*out) & ~s->want)
    return fail(c, STG_RT_EINVAL, "set_download: array not part of the set's want mask");
  const uint64_t n = s->n_beams;
  if (out->ranges) CU(c, cudaMemcpyAsync(out->ranges, s->d.ranges.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (out->intensities) CU(c, cudaMemcpyAsync(out->intensities, s->d.intens.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (out->bearings) CU(c, cudaMemcpyAsync(out->bearings, s->d.bearings.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (out->hit_model) CU(c, cudaMemcpyAsync(out->hit_model, s->d.hit_model.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (out->hit_cell) CU(c, cudaMemcpyAsync(out->hit_cell, s->d.hit_cell.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
  if (out->steps) CU(c, cudaMemcpyAsync(out->steps, s->d.steps.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return STG_RT_OK;
}

int stg_rt_set_device_ptrs(stg_rt_ctx *c, stg_rt_set *s, stg_rt_fan_out *dev)
{
  if (!c || !s || !dev)
    return STG_RT_EINVAL;
  const FanBatchView v = view_of(s->d, (uint32_t)s->fans.size(), s->n_beams, s->want, false, false);
  dev->ranges = v.ranges;
  dev->intensities = v.intensities;
  dev->bearings = v.bearings;
  dev->hit_model = v.hit_model;
  dev->hit_cell = v.hit_cell;
  dev->steps = v.steps;
  return STG_RT_OK;
}

uint64_t stg_rt_set_beams(const stg_rt_set *s) { return s ? s->n_beams : 0; }

// ---- grid inspection ------------------------------------------------------------------------------

int64_t stg_rt_grid_dump(stg_rt_ctx *c, int32_t *records, int64_t cap, int64_t *region_counts, int64_t counts_cap,
                         int64_t *n_counts)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc)
    return rc;
  if ((rc = deliver_locked(c)))
    return rc;
  const GridView &g = c->g;
  std::vector<unsigned long long> slots(c->slot_cap);
  std::vector<BlockRec> recs_dev(c->cfg.max_blocks);
  std::vector<uint32_t> counts(c->n_regions);
  struct Rec {
    int32_t layer, x, y;
    unsigned long long seq;
    int32_t block;
  };
  std::vector<Rec> recs;
  for (int L = 0; L < 2; L++) {
    CU(c, cudaMemcpyAsync(slots.data(), g.slots[L], (size_t)c->slot_cap * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(recs_dev.data(), g.blk_rec[L], (size_t)c->cfg.max_blocks * sizeof(BlockRec), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < slots.size(); i++) {
      const unsigned long long e = slots[i];
      if (e == kSlotEmpty || e == kSlotTomb)
        continue;
      const uint32_t key = (uint32_t)(e >> 32), blk = (uint32_t)e - 1u;
      const uint32_t region = key >> 10, cy = (key >> 5) & 31, cx = key & 31;
      Rec r;
      r.layer = L;
      r.x = (g.rx0 + (int32_t)(region % (uint32_t)g.nrx)) * 32 + (int32_t)cx;
      r.y = (g.ry0 + (int32_t)(region / (uint32_t)g.nrx)) * 32 + (int32_t)cy;
      r.seq = blk < recs_dev.size() ? recs_dev[blk].seq : 0;
      r.block = (int32_t)blk;
      recs.push_back(r);
    }
  }
  std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {
    if (a.layer != b.layer) return a.layer < b.layer;
    if (a.y != b.y) return a.y < b.y;
    if (a.x != b.x) return a.x < b.x;
    if (a.seq != b.seq) return a.seq < b.seq;
    return a.block < b.block;
  });
  int64_t n = 0;
  int32_t pos = 0;
  for (size_t i = 0; i < recs.size(); i++) {
    pos = (i && recs[i - 1].layer == recs[i].layer && recs[i - 1].x == recs[i].x && recs[i - 1].y == recs[i].y) ? pos + 1 : 0;
    if (records && n < cap) {
      int32_t *o = records + 5 * n;
      o[0] = recs[i].layer;
      o[1] = recs[i].x;
      o[2] = recs[i].y;
      o[3] = pos;
      o[4] = recs[i].block;
    }
    n++;
  }
  CU(c, cudaMemcpy2DAsync(counts.data(), 4, &g.head[0].count, sizeof(RegionHead), 4, c->n_regions, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  int64_t nc = 0;
  for (size_t r = 0; r < counts.size(); r++)
    if (counts[r]) {
      if (region_counts && nc < counts_cap) {
        region_counts[3 * nc] = g.rx0 + (int64_t)(r % (size_t)g.nrx);
        region_counts[3 * nc + 1] = g.ry0 + (int64_t)(r / (size_t)g.nrx);
        region_counts[3 * nc + 2] = counts[r];
      }
      nc++;
    }
  if (n_counts)
    *n_counts = nc;
  return n;
}

int stg_rt_grid_hash(stg_rt_ctx *c, uint64_t out[5])
{
  if (!c || !out)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc || (rc = deliver_locked(c)))
    return rc;
  DevBuf d;
  if ((rc = dev_reserve(c, d, 5 * 8)))
    return rc;
  cudaError_t e = cudaMemsetAsync(d.p, 0, 5 * 8, c->stream);
  if (e == cudaSuccess) {
    launch_grid_hash(c->g, (unsigned long long *)d.p, c->stream);
    c->stats.launches_other++;
    e = cudaMemcpyAsync(out, d.p, 5 * 8, cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  cudaFree(d.p);
  if (e != cudaSuccess)
    return fail(c, STG_RT_ECUDA, "grid_hash: %s", cudaGetErrorString(e));
  return STG_RT_OK;
}

int64_t stg_rt_debug_region_owners(stg_rt_ctx *c, int64_t *records, int64_t cap)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc)
    return rc;
  if ((rc = deliver_locked(c)))
    return rc;
  const GridView &g = c->g;
  std::vector<uint32_t> owners(c->n_regions);
  CU(c, cudaMemcpy2DAsync(owners.data(), 4, &g.head[0].owner, sizeof(RegionHead), 4, c->n_regions, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  int64_t n = 0;
  for (size_t r = 0; r < owners.size(); r++)
    if (owners[r]) {
      if (records && n < cap) {
        records[3 * n] = g.rx0 + (int64_t)(r % (size_t)g.nrx);
        records[3 * n + 1] = g.ry0 + (int64_t)(r / (size_t)g.nrx);
        records[3 * n + 2] = owners[r];
      }
      n++;
    }
  return n;
}

int stg_rt_debug_occupancy_check(stg_rt_ctx *c, uint64_t out[3])
{
  if (!c || !out)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc)
    return rc;
  if ((rc = deliver_locked(c)))
    return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaStreamSynchronize(c->sum_stream)); // k1_summaries of the last pass has run: the summaries are exact now
  const GridView &g = c->g;
  const size_t nr = c->n_regions;
  std::vector<uint32_t> occ(occ_alloc_words(nr)), want(2 * nr * kTileWords, 0u);
  std::vector<RegionHead> heads(nr);
  std::vector<unsigned long long> slots(c->slot_cap);
  CU(c, cudaMemcpyAsync(occ.data(), g.occ[0], occ.size() * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(heads.data(), g.head, nr * sizeof(RegionHead), cudaMemcpyDeviceToHost, c->stream));
  for (int L = 0; L < 2; L++) {
    CU(c, cudaMemcpyAsync(slots.data(), g.slots[L], (size_t)c->slot_cap * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    uint32_t *w = want.data() + (size_t)L * nr * kTileWords;
    for (const unsigned long long e : slots) {
      if (e == kSlotEmpty || e == kSlotTomb)
        continue;
      const uint32_t key = (uint32_t)(e >> 32), region = key >> 10, cy = (key >> 5) & 31, cx = key & 31;
      if (region >= nr)
        continue;
      w[(size_t)region * kTileWords + cy] |= 1u << cx;
      w[(size_t)region * kTileWords + kRegionWidth + cx] |= 1u << cy;
    }
  }
  out[0] = out[1] = out[2] = 0;
  for (int L = 0; L < 2; L++)
    for (size_t r = 0; r < nr; r++) {
      const uint32_t *have = occ.data() + ((size_t)L * nr + r) * kTileWords, *w = want.data() + ((size_t)L * nr + r) * kTileWords;
      const uint32_t *sum = heads[r].sum[L];
      uint32_t rows = 0, cols = 0;
      for (int i = 0; i < kRegionWidth; i++) {
        out[0] += (uint64_t)__builtin_popcount(have[i] ^ w[i]) + (uint64_t)__builtin_popcount(have[kRegionWidth + i] ^ w[kRegionWidth + i]);
        out[2] += (uint64_t)__builtin_popcount(w[i]);
        rows |= have[i] ? 1u << i : 0u;
        cols |= have[kRegionWidth + i] ? 1u << i : 0u;
      }
      out[1] += (uint64_t)__builtin_popcount(rows ^ sum[0]) + (uint64_t)__builtin_popcount(cols ^ sum[1]);
    }
  for (size_t i = 0; i < (nr + 31) / 32; i++) // a dirty bit left standing is a summary nobody will bring up to date
    out[1] += (uint64_t)__builtin_popcount(occ[2 * nr * kTileWords + i]);
  return STG_RT_OK;
}

// ---- multi-GPU deltas ------------------------------------------------------------------------------

size_t stg_rt_delta_slot_bytes(const stg_rt_ctx *c) { return c ? c->blob_cap : 0; }

int stg_rt_delta_export(stg_rt_ctx *c, int rank, void **dev_blob, size_t *n_bytes)
{
  if (!c || !dev_blob || !n_bytes || rank < 0 || rank > 255)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = launch_fans_locked(c); // fans queued so far see the grid before these deltas
  if (rc)
    return rc;
  c->last_export_rank = rank;
  c->export_mode = true;
  // the entry accounting of exported deltas happens here for this rank's own share
  if ((rc = build_and_send_deltas(c, false, rank, n_bytes)))
    return rc;
  *dev_blob = c->d_blob.p;
  return STG_RT_OK;
}

int stg_rt_stream_signal(stg_rt_ctx *c, void *consumer_stream)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if ((cudaStream_t)consumer_stream == c->stream)
    return STG_RT_OK;
  CU(c, cudaEventRecord(c->xstream_ev, c->stream));
  CU(c, cudaStreamWaitEvent((cudaStream_t)consumer_stream, c->xstream_ev, 0));
  return STG_RT_OK;
}

int stg_rt_stream_wait(stg_rt_ctx *c, void *producer_stream)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if ((cudaStream_t)producer_stream == c->stream)
    return STG_RT_OK;
  CU(c, cudaEventRecord(c->xstream_ev, (cudaStream_t)producer_stream));
  CU(c, cudaStreamWaitEvent(c->stream, c->xstream_ev, 0));
  return STG_RT_OK;
}

int stg_rt_delta_apply_gathered(stg_rt_ctx *c, const void *dev_gathered, int n_ranks, size_t slot_bytes)
{
  if (!c || !dev_gathered || n_ranks < 1 || n_ranks > 256 || slot_bytes < sizeof(DeltaHeader))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = launch_fans_locked(c);
  if (rc)
    return rc;
  // Table accounting needs what the OTHER ranks inserted / removed, which only their headers (now on
  // the device) know. They are read back asynchronously and booked one exchange late, so the host
  // never waits for the GPU here; until then this tick is charged like the last one (the compaction
  // threshold of 3/4 and the half-empty rule leave ample slack for a tick's worth of error; an insert that
  // finds no slot all the same raises kErrTableFull on the device and fails the next stg_rt_sync).
  if (c->hdrs_pending) {
    CU(c, cudaEventSynchronize(c->hdrs_ready));
    const DeltaHeader *h = (const DeltaHeader *)c->h_hdrs.p;
    int64_t all_ins[2] = { 0, 0 }, all_rem[2] = { 0, 0 };
    for (int r = 0; r < c->hdrs_ranks; r++) {
      if (h[r].magic != kDeltaMagic)
        continue; // K1 skipped it and said so
      c->epoch = std::max<uint64_t>(c->epoch, h[r].epoch); // the epoch that pass really used
      for (int L = 0; L < 2; L++) {
        all_ins[L] += (int64_t)h[r].steps_insert_l[L] + h[r].est_insert_l[L];
        all_rem[L] += (int64_t)h[r].steps_remove_l[L] + h[r].est_remove_l[L];
        if ((int)h[r].rank != c->last_export_rank) // this rank's own share was booked at export; the others' net change, signed
          c->live_entries[L] += (int64_t)h[r].steps_insert_l[L] + h[r].est_insert_l[L] - (int64_t)h[r].steps_remove_l[L] - h[r].est_remove_l[L];
      }
    }
    for (int L = 0; L < 2; L++) {
      c->last_gather_ins[L] = all_ins[L];
      c->last_gather_rem[L] = all_rem[L];
    }
    c->hdrs_pending = false;
  }
  if ((rc = upload_geometry_locked(c)))
    return rc;
  int rcp = pin_reserve(c, c->h_hdrs, (size_t)n_ranks * sizeof(DeltaHeader));
  if (rcp)
    return rcp;
  CU(c, cudaMemcpy2DAsync(c->h_hdrs.p, sizeof(DeltaHeader), dev_gathered, slot_bytes, sizeof(DeltaHeader), n_ranks,
                          cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaEventRecord(c->hdrs_ready, c->stream));
  c->hdrs_pending = true;
  c->hdrs_ranks = n_ranks;
  int64_t add[2];
  for (int L = 0; L < 2; L++) {
    add[L] = std::max<int64_t>(c->last_gather_ins[L], (int64_t)n_ranks * c->last_export_ins[L]);
    // what this exchange is expected to leave in the table: as the last one did (every rank's movers re-rendered: as much
    // goes as comes), or this rank's own share times the ranks before anything is known
    const int64_t going = c->last_gather_ins[L] ? c->last_gather_rem[L] : (int64_t)n_ranks * c->last_export_rem[L];
    if (c->live_entries[L] + std::max<int64_t>(add[L] - going, 0) > (int64_t)c->slot_cap / 2)
      return fail(c, STG_RT_ENOMEM, "cell-entry table of layer %d full (%lld entries, %lld arriving, %lld leaving, capacity %u/2); raise max_cell_entries",
                  L, (long long)c->live_entries[L], (long long)add[L], (long long)going, c->slot_cap);
  }
  c->epoch++; // K1 takes the largest proposal among the blobs, which is at least this (every sender proposed its own epoch + 1)
  if ((rc = summaries_fence_locked(c)))
    return rc;
  CU(c, cudaEventRecord(c->t_ev[0][0], c->stream));
  launch_apply_deltas(c->g, c->gv, (const unsigned char *)dev_gathered, n_ranks, slot_bytes, c->epoch, 0, kDeltaEdgeRecords | kDeltaPoseRecords, c->stream,
                      &c->stats.launches_rasterise);
  if ((rc = ensure_table_room(c, add)))
    return rc;
  launch_apply_deltas(c->g, c->gv, (const unsigned char *)dev_gathered, n_ranks, slot_bytes, c->epoch, 1, kDeltaEdgeRecords | kDeltaPoseRecords, c->stream,
                      &c->stats.launches_rasterise);
  launch_owner_rebuild(c->g, c->stream, &c->stats.launches_rasterise); // other ranks' updates are only known there
  CU(c, cudaEventRecord(c->t_ev[0][1], c->stream));
  if ((rc = summaries_launch_locked(c)))
    return rc;
  c->t_rec[0] = true;
  CU(c, cudaGetLastError());
  return STG_RT_OK;
}

// ---- sensor post-processing: fiducial finders and blobfinders ---------------------------------------

static int fiducials_submit_impl(stg_rt_ctx *c, uint32_t n_targets, const stg_rt_fid_target *targets, uint32_t n_finders,
                                 const stg_rt_fid_finder *finders, uint32_t max_per_finder, stg_rt_fiducial *out, uint32_t *out_count,
                                 uint64_t packed_capacity, const void *dev_targets = nullptr);

int stg_rt_fiducials_submit(stg_rt_ctx *c, uint32_t n_targets, const stg_rt_fid_target *targets, uint32_t n_finders,
                            const stg_rt_fid_finder *finders, uint32_t max_per_finder, stg_rt_fiducial *out, uint32_t *out_count)
{
  return fiducials_submit_impl(c, n_targets, targets, n_finders, finders, max_per_finder, out, out_count, 0);
}

int stg_rt_fiducials_submit_packed(stg_rt_ctx *c, uint32_t n_targets, const stg_rt_fid_target *targets, uint32_t n_finders,
                                   const stg_rt_fid_finder *finders, uint32_t max_per_finder, uint64_t out_capacity, stg_rt_fiducial *out,
                                   uint32_t *out_count)
{
  if (!out_capacity)
    return STG_RT_EINVAL;
  return fiducials_submit_impl(c, n_targets, targets, n_finders, finders, max_per_finder, out, out_count, out_capacity);
}

int stg_rt_fiducials_submit_device_targets(stg_rt_ctx *c, uint32_t n_targets, const void *dev_targets, uint32_t n_finders,
                                           const stg_rt_fid_finder *finders, uint32_t max_per_finder, uint64_t out_capacity, stg_rt_fiducial *out,
                                           uint32_t *out_count)
{
  if (!out_capacity || (n_targets && !dev_targets))
    return STG_RT_EINVAL;
  return fiducials_submit_impl(c, n_targets, nullptr, n_finders, finders, max_per_finder, out, out_count, out_capacity, dev_targets);
}

static int fiducials_submit_impl(stg_rt_ctx *c, uint32_t n_targets, const stg_rt_fid_target *targets, uint32_t n_finders,
                                 const stg_rt_fid_finder *finders, uint32_t max_per_finder, stg_rt_fiducial *out, uint32_t *out_count,
                                 uint64_t packed_capacity, const void *dev_targets)
{
  static_assert(sizeof(stg_rt_fid_target) == sizeof(FidTargetRec) && sizeof(stg_rt_fid_finder) == sizeof(FidFinderRec) &&
                    sizeof(stg_rt_fiducial) == sizeof(FiducialRec), "fiducial record layouts");
  if (!c || (n_targets && !targets && !dev_targets) || (n_finders && (!finders || !out || !out_count)) || !max_per_finder)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!n_finders)
    return STG_RT_OK;
  for (uint32_t i = 0; i < n_finders; i++)
    if (finders[i].finder >= c->cfg.max_models || !c->model_known[finders[i].finder] || finders[i].layer > 1)
      return fail(c, STG_RT_EINVAL, "fiducial finder %u: unknown model %u or bad layer", i, finders[i].finder);
  for (uint32_t i = 0; i < n_targets && targets; i++)
    if (targets[i].model >= c->cfg.max_models || !c->model_known[targets[i].model])
      return fail(c, STG_RT_EINVAL, "fiducial target %u: unknown model %u", i, targets[i].model);
  if (dev_targets && (c->cfg.flags & STG_RT_CFG_HOST_EXACT_FIDUCIALS))
    return fail(c, STG_RT_ESTATE, "fiducials_submit_device_targets: STG_RT_CFG_HOST_EXACT_FIDUCIALS needs the targets on the host");
  int rc = flush_locked(c); // issue order: everything submitted before is on the stream first
  if (rc || (rc = deliver_post_locked(c)))
    return rc;
  // targets already on the device (the output of the ranks' all-gather) are read where they are; only the finders go up
  const size_t b_t = dev_targets ? 0 : (size_t)n_targets * sizeof(FidTargetRec), b_f = (size_t)n_finders * sizeof(FidFinderRec);
  const size_t b_fan = (size_t)n_finders * sizeof(FanRec), b_in = b_t + b_f + b_fan;
  const size_t b_out = (size_t)n_finders * max_per_finder * sizeof(FiducialRec), b_cnt = (size_t)n_finders * 4;
  if ((rc = pin_reserve(c, c->h_fid_in, b_in)) || (rc = dev_reserve(c, c->d_fid_in, b_in)) ||
      (rc = pin_reserve(c, c->h_fid_out, b_out)) || (rc = dev_reserve(c, c->d_fid_out, b_out)) ||
      (rc = pin_reserve(c, c->h_fid_count, b_cnt)) || (rc = dev_reserve(c, c->d_fid_count, b_cnt)))
    return rc;
  char *hp = (char *)c->h_fid_in.p;
  // staged by the host threads side by side (100,000 finders are 10 MB; one core copies them in a millisecond)
  // (inline when they are busy expanding the intensities of a march still running)
  c->pool.parallel_for(b_t + b_f, c->pool.busy() ? b_t + b_f + 1 : (size_t)1 << 20, [hp, targets, finders, b_t](size_t lo, size_t hi) {
    if (lo < b_t)
      memcpy(hp + lo, (const char *)targets + lo, std::min(hi, b_t) - lo);
    if (hi > b_t) {
      const size_t from = std::max(lo, b_t);
      memcpy(hp + from, (const char *)finders + (from - b_t), hi - from);
    }
  });
  CU(c, cudaMemcpyAsync(c->d_fid_in.p, hp, b_t + b_f, cudaMemcpyHostToDevice, c->stream));
  launch_fid_fans((const FidFinderRec *)((char *)c->d_fid_in.p + b_t), (FanRec *)((char *)c->d_fid_in.p + b_t + b_f), n_finders, c->stream);
  c->stats.launches_post++;
  c->stats.h2d_bytes += b_t + b_f;
  FidBatchView v;
  v.n_targets = n_targets;
  v.n_finders = n_finders;
  v.max_per_finder = max_per_finder;
  v.targets = dev_targets ? (const FidTargetRec *)dev_targets : (const FidTargetRec *)c->d_fid_in.p;
  v.finders = (const FidFinderRec *)((char *)c->d_fid_in.p + b_t);
  v.fans = (const FanRec *)((char *)c->d_fid_in.p + b_t + b_f);
  v.out = (FiducialRec *)c->d_fid_out.p;
  v.out_count = (uint32_t *)c->d_fid_count.p;
  // Candidate query: brute force over the target list for small problems, a per-submit uniform grid
  // (bins a quarter of the longest detection range wide) beyond ~4 M finder x target pairs.
  // STG_RT_FID_GRID=0/1 forces one or the other (tests compare the two).
  CU(c, cudaEventRecord(c->t_ev[3][0], c->stream));
  const char *force = getenv("STG_RT_FID_GRID");
  const bool use_grid = force ? atoi(force) != 0 : (uint64_t)n_targets * n_finders > (4u << 20);
  if (use_grid && n_targets) {
    // the bins cover the targets' bounding box - or, when the targets are only on the device, the grid's extent
    // (targets outside it fall into the border bins, which the finders' windows reach like any other)
    double x0 = c->cell_x0 / c->cfg.ppm, x1 = (c->cell_x1 + 1) / c->cfg.ppm, y0 = c->cell_y0 / c->cfg.ppm, y1 = (c->cell_y1 + 1) / c->cfg.ppm, cell = 0;
    if (targets) {
      x0 = x1 = targets[0].x;
      y0 = y1 = targets[0].y;
    }
    for (uint32_t i = 1; i < n_targets && targets; i++) {
      x0 = std::min(x0, targets[i].x);
      x1 = std::max(x1, targets[i].x);
      y0 = std::min(y0, targets[i].y);
      y1 = std::max(y1, targets[i].y);
    }
    for (uint32_t i = 0; i < n_finders; i++)
      cell = std::max(cell, finders[i].max_range_anon);
    if (!(cell > 0))
      cell = 1.0;
    cell /= 4; // a finder's window is at most 9 rows of bins; each row is sifted only where it meets the finder's (half) disc
    while ((x1 - x0) / cell * ((y1 - y0) / cell) > 4.0e6) // keep the bin table small
      cell *= 2;
    FidGridView gv;
    gv.x0 = x0;
    gv.y0 = y0;
    gv.cell = cell;
    gv.nbx = (int32_t)floor((x1 - x0) / cell) + 1;
    gv.nby = (int32_t)floor((y1 - y0) / cell) + 1;
    const size_t n_bins = (size_t)gv.nbx * gv.nby;
    const size_t b_grid = ((size_t)n_targets * 2 + n_bins * 2 + 1) * 4;
    if ((rc = dev_reserve(c, c->d_fid_grid, b_grid)))
      return rc;
    gv.bin_of = (uint32_t *)c->d_fid_grid.p;
    gv.sorted = gv.bin_of + n_targets;
    gv.start = gv.sorted + n_targets;
    gv.cursor = gv.start + n_bins + 1;
    CU(c, cudaMemsetAsync(gv.start, 0, (n_bins + 1) * 4, c->stream));
    launch_fiducials_grid(c->g, v, gv, c->stream);
    c->stats.launches_post += 4;
  } else {
    launch_fiducials(c->g, v, c->stream);
    c->stats.launches_post++;
  }
  // only the records come home (packed, written by the device into the page-locked staging buffer);
  // stg_rt_sync puts them back at the caller's fixed stride
  if ((rc = dev_reserve(c, c->d_fid_off, (size_t)(n_finders + 1) * 4)))
    return rc;
  launch_fid_pack(v, (uint32_t *)c->d_fid_off.p, c->h_fid_out.p, (uint32_t *)c->h_fid_count.p, c->stream);
  CU(c, cudaEventRecord(c->t_ev[3][1], c->stream));
  c->t_rec[3] = true;
  c->stats.launches_post += 2;
  CU(c, cudaGetLastError());
  c->fid_job.user_out = out;
  c->fid_job.user_count = out_count;
  c->fid_job.n_finders = n_finders;
  c->fid_job.n_targets = n_targets;
  c->fid_job.max_per_finder = max_per_finder;
  c->fid_job.packed_capacity = packed_capacity;
  c->fid_job.pending = true;
  return STG_RT_OK;
}

int stg_rt_collisions_test(stg_rt_ctx *c, int layer, uint32_t n_queries, const uint32_t *query_first, const uint32_t *blocks,
                           uint32_t ground_model, int32_t *hit_model)
{
  if (!c || layer < 0 || layer > 1 || (n_queries && (!query_first || !hit_model)))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!n_queries)
    return STG_RT_OK;
  const uint32_t n_ids = query_first[n_queries];
  if (n_ids && !blocks)
    return STG_RT_EINVAL;
  for (uint32_t q = 0; q < n_queries; q++) {
    if (query_first[q + 1] < query_first[q])
      return fail(c, STG_RT_EINVAL, "collisions_test: query_first is not ascending at %u", q);
    for (uint32_t i = query_first[q]; i < query_first[q + 1]; i++)
      if (blocks[i] >= c->cfg.max_blocks || !c->blocks[blocks[i]].known)
        return fail(c, STG_RT_EINVAL, "collisions_test: query %u names unknown block %u", q, blocks[i]);
  }
  int rc = flush_locked(c); // the grid the test reads includes every Map / UnMap issued so far
  if (rc || (rc = deliver_post_locked(c)))
    return rc;
  size_t n_edges = 0;
  for (uint32_t i = 0; i < n_ids; i++) {
    if (c->blocks[blocks[i]].dev_mapped[layer] && c->blocks[blocks[i]].dev_owned[layer] && (rc = fetch_outline_locked(c, blocks[i], layer)))
      return rc; // rendered by the device from a pose: the walk below needs the outline
    n_edges += c->blocks[blocks[i]].dev_xy[layer].size() / 2;
  }
  const size_t b_q = (size_t)n_queries * sizeof(CollisionQuery), b_in = b_q + n_edges * sizeof(EdgeRec);
  if ((rc = pin_reserve(c, c->h_col_in, b_in)) || (rc = dev_reserve(c, c->d_col_in, b_in + 16)) ||
      (rc = pin_reserve(c, c->h_col_hit, (size_t)n_queries * 4)) || (rc = dev_reserve(c, c->d_col_hit, (size_t)n_queries * 4)))
    return rc;
  CollisionQuery *qr = (CollisionQuery *)c->h_col_in.p;
  EdgeRec *e0 = (EdgeRec *)((char *)c->h_col_in.p + b_q), *e = e0;
  c->col_job.fallback.assign(n_queries, -1);
  for (uint32_t q = 0; q < n_queries; q++) {
    qr[q].first_edge = (uint32_t)(e - e0);
    qr[q].layer = (uint32_t)layer;
    uint32_t cursor = 0;
    for (uint32_t i = query_first[q]; i < query_first[q + 1]; i++) {
      const BlockShadow &b = c->blocks[blocks[i]];
      if (!(c->models[b.model].flags & STG_RT_VIS_OBSTACLE_RETURN)) // block.cc:136
        continue;
      if (b.zmin < 0) { // block.cc:137-138: the ground, unless an earlier block already hit something
        c->col_job.fallback[q] = (int32_t)ground_model;
        break;
      }
      if (b.dev_mapped[layer]) // after the flush dev_xy is the outline rendered on the device
        emit_edges(b.dev_xy[layer], blocks[i], (uint32_t)layer, cursor, e, &b.dev_poly[layer]);
    }
    qr[q].n_edges = (uint32_t)(e - e0) - qr[q].first_edge;
    qr[q].n_steps = cursor;
  }
  const size_t b_used = b_q + (size_t)(e - e0) * sizeof(EdgeRec);
  CU(c, cudaMemcpyAsync(c->d_col_in.p, c->h_col_in.p, b_used, cudaMemcpyHostToDevice, c->stream));
  c->stats.h2d_bytes += b_used;
  CollisionBatchView v;
  v.n_queries = n_queries;
  v.queries = (const CollisionQuery *)c->d_col_in.p;
  v.edges = (const EdgeRec *)((const char *)c->d_col_in.p + b_q);
  v.hit_model = (int32_t *)c->d_col_hit.p;
  CU(c, cudaEventRecord(c->t_ev[5][0], c->stream));
  launch_collisions(c->g, v, c->stream);
  CU(c, cudaEventRecord(c->t_ev[5][1], c->stream));
  c->t_rec[5] = true;
  c->stats.launches_post++;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(c->h_col_hit.p, c->d_col_hit.p, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, c->stream));
  c->stats.d2h_bytes += (size_t)n_queries * 4;
  c->col_job.user_hit = hit_model;
  c->col_job.pending = true;
  return STG_RT_OK;
}

int stg_rt_touching_models(stg_rt_ctx *c, int layer, uint32_t n_queries, const uint32_t *query_first, const uint32_t *blocks,
                           uint32_t max_per_query, uint32_t *touchers, uint32_t *counts)
{
  if (!c || layer < 0 || layer > 1 || (n_queries && (!query_first || !counts || (max_per_query && !touchers))))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!n_queries)
    return STG_RT_OK;
  const uint32_t n_ids = query_first[n_queries];
  if (n_ids && !blocks)
    return STG_RT_EINVAL;
  for (uint32_t q = 0; q < n_queries; q++) {
    if (query_first[q + 1] < query_first[q])
      return fail(c, STG_RT_EINVAL, "touching_models: query_first is not ascending at %u", q);
    for (uint32_t i = query_first[q]; i < query_first[q + 1]; i++)
      if (blocks[i] >= c->cfg.max_blocks || !c->blocks[blocks[i]].known)
        return fail(c, STG_RT_EINVAL, "touching_models: query %u names unknown block %u", q, blocks[i]);
  }
  int rc = flush_locked(c); // the grid the walk reads includes every Map / UnMap issued so far
  if (rc || (rc = deliver_post_locked(c)))
    return rc;
  size_t n_edges = 0;
  for (uint32_t i = 0; i < n_ids; i++) {
    if (c->blocks[blocks[i]].dev_mapped[layer] && c->blocks[blocks[i]].dev_owned[layer] && (rc = fetch_outline_locked(c, blocks[i], layer)))
      return rc;
    n_edges += c->blocks[blocks[i]].dev_xy[layer].size() / 2;
  }
  const size_t b_q = (size_t)n_queries * sizeof(CollisionQuery), b_in = b_q + n_edges * sizeof(EdgeRec);
  const size_t b_out = (size_t)n_queries * 4 * (1 + (size_t)max_per_query);
  if ((rc = pin_reserve(c, c->h_col_in, b_in)) || (rc = dev_reserve(c, c->d_col_in, b_in + 16)) ||
      (rc = pin_reserve(c, c->h_touch_out, b_out)) || (rc = dev_reserve(c, c->d_touch_out, b_out)))
    return rc;
  CollisionQuery *qr = (CollisionQuery *)c->h_col_in.p;
  EdgeRec *e0 = (EdgeRec *)((char *)c->h_col_in.p + b_q), *e = e0;
  for (uint32_t q = 0; q < n_queries; q++) {
    qr[q].first_edge = (uint32_t)(e - e0);
    qr[q].layer = (uint32_t)layer;
    uint32_t cursor = 0;
    for (uint32_t i = query_first[q]; i < query_first[q + 1]; i++) {
      const BlockShadow &b = c->blocks[blocks[i]];
      if (b.dev_mapped[layer]) // Block::rendered_cells[layer]: what the block is rendered into now
        emit_edges(b.dev_xy[layer], blocks[i], (uint32_t)layer, cursor, e, &b.dev_poly[layer]);
    }
    qr[q].n_edges = (uint32_t)(e - e0) - qr[q].first_edge;
    qr[q].n_steps = cursor;
  }
  const size_t b_used = b_q + (size_t)(e - e0) * sizeof(EdgeRec);
  CU(c, cudaMemcpyAsync(c->d_col_in.p, c->h_col_in.p, b_used, cudaMemcpyHostToDevice, c->stream));
  c->stats.h2d_bytes += b_used;
  ToucherBatchView v;
  v.n_queries = n_queries;
  v.max_per_query = max_per_query;
  v.queries = (const CollisionQuery *)c->d_col_in.p;
  v.edges = (const EdgeRec *)((const char *)c->d_col_in.p + b_q);
  v.out_count = (uint32_t *)c->d_touch_out.p;
  v.out = v.out_count + n_queries;
  CU(c, cudaEventRecord(c->t_ev[5][0], c->stream));
  launch_touchers(c->g, v, c->stream);
  CU(c, cudaEventRecord(c->t_ev[5][1], c->stream));
  c->t_rec[5] = true;
  c->stats.launches_post++;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(c->h_touch_out.p, c->d_touch_out.p, b_out, cudaMemcpyDeviceToHost, c->stream));
  c->stats.d2h_bytes += b_out;
  c->touch_job.user_out = touchers;
  c->touch_job.user_count = counts;
  c->touch_job.n_queries = n_queries;
  c->touch_job.max_per_query = max_per_query;
  c->touch_job.pending = true;
  return STG_RT_OK;
}

int stg_rt_models_move(stg_rt_ctx *c, int layer, uint32_t n_movers, const uint32_t *mover_first, const uint32_t *blocks,
                       const double *z, const uint32_t *first_pt, const int32_t *xy, const double *z_start,
                       const uint32_t *first_pt_start, const int32_t *xy_start, uint8_t *stalled)
{
  if (!c || layer < 0 || layer > 1 ||
      (n_movers && (!mover_first || !blocks || !z || !first_pt || !xy || !z_start || !first_pt_start || !xy_start || !stalled)))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!n_movers)
    return STG_RT_OK;
  size_t n_edges[2] = { 0, 0 }; // of the new outlines, of the start outlines
  std::vector<uint32_t> pairs;  // (model, mover index + 1)
  for (uint32_t i = 0; i < n_movers; i++) {
    if (mover_first[i + 1] < mover_first[i])
      return fail(c, STG_RT_EINVAL, "models_move: mover_first is not ascending at %u", i);
    for (uint32_t k = mover_first[i]; k < mover_first[i + 1]; k++) {
      if (blocks[k] >= c->cfg.max_blocks || !c->blocks[blocks[k]].known)
        return fail(c, STG_RT_EINVAL, "models_move: mover %u names unknown block %u", i, blocks[k]);
      if (first_pt[k + 1] < first_pt[k] || first_pt_start[k + 1] < first_pt_start[k])
        return fail(c, STG_RT_EINVAL, "models_move: first_pt is not ascending at %u", k);
      n_edges[0] += first_pt[k + 1] - first_pt[k];
      n_edges[1] += first_pt_start[k + 1] - first_pt_start[k];
      const uint32_t model = c->blocks[blocks[k]].model;
      if (pairs.size() < 2 || pairs[pairs.size() - 2] != model) {
        pairs.push_back(model);
        pairs.push_back(i + 1);
      }
    }
  }
  int rc = flush_locked(c); // the grid as it stands before the first mover moves
  if (rc || (rc = deliver_locked(c)))
    return rc;
  // ---- two sets of queries: the outlines the movers want, and the outlines they have now. Blocks of models that
  // are no obstacles test nothing and are tested by nobody (block.cc:136,154).
  const size_t b_q = (size_t)n_movers * sizeof(CollisionQuery);
  size_t off = 0, o_q[2], o_e[2], o_z[2];
  for (int s = 0; s < 2; s++) {
    o_q[s] = off;
    off += b_q;
    o_e[s] = off;
    off += n_edges[s] * sizeof(EdgeRec);
    o_z[s] = off;
    off += n_edges[s] * 16;
  }
  const size_t o_p = off, b_p = pairs.size() * 4, b_in = o_p + b_p;
  const size_t b_out = (size_t)n_movers * sizeof(TouchRec);
  if ((rc = pin_reserve(c, c->h_mv_in, b_in)) || (rc = dev_reserve(c, c->d_mv_in, b_in + 16)) ||
      (rc = pin_reserve(c, c->h_mv_out, 3 * b_out)) || (rc = dev_reserve(c, c->d_mv_out, 3 * b_out)))
    return rc;
  if (!c->d_mover_of_model.p) {
    if ((rc = dev_reserve(c, c->d_mover_of_model, (size_t)c->cfg.max_models * 4)))
      return rc;
    CU(c, cudaMemsetAsync(c->d_mover_of_model.p, 0, (size_t)c->cfg.max_models * 4, c->stream));
  }
  char *hp = (char *)c->h_mv_in.p;
  std::vector<uint8_t> ground(n_movers, 0); // a block of an obstacle model reaching below the floor: block.cc:137-138
  std::vector<int32_t> tmp_xy;
  for (int s = 0; s < 2; s++) {
    const double *zz = s ? z_start : z;
    const uint32_t *fp = s ? first_pt_start : first_pt;
    const int32_t *pts = s ? xy_start : xy;
    CollisionQuery *qr = (CollisionQuery *)(hp + o_q[s]);
    EdgeRec *e0 = (EdgeRec *)(hp + o_e[s]), *e = e0;
    double *ez = (double *)(hp + o_z[s]);
    for (uint32_t i = 0; i < n_movers; i++) {
      qr[i].first_edge = (uint32_t)(e - e0);
      qr[i].layer = (uint32_t)layer;
      uint32_t cursor = 0;
      for (uint32_t k = mover_first[i]; k < mover_first[i + 1]; k++) {
        const BlockShadow &b = c->blocks[blocks[k]];
        if (!(c->models[b.model].flags & STG_RT_VIS_OBSTACLE_RETURN))
          continue;
        if (s == 0 && zz[2 * k] < 0)
          ground[i] = 1;
        tmp_xy.assign(pts + 2 * (size_t)fp[k], pts + 2 * (size_t)fp[k + 1]);
        const EdgeRec *before = e;
        emit_edges(tmp_xy, blocks[k], (uint32_t)layer, cursor, e);
        for (const EdgeRec *p = before; p < e; p++) {
          ez[2 * (p - e0)] = zz[2 * k];
          ez[2 * (p - e0) + 1] = zz[2 * k + 1];
        }
      }
      qr[i].n_edges = (uint32_t)(e - e0) - qr[i].first_edge;
      qr[i].n_steps = cursor;
    }
  }
  memcpy(hp + o_p, pairs.data(), b_p);
  CU(c, cudaMemcpyAsync(c->d_mv_in.p, hp, b_in, cudaMemcpyHostToDevice, c->stream));
  c->stats.h2d_bytes += b_in;
  const char *dp = (const char *)c->d_mv_in.p;
  const uint32_t *d_pairs = (const uint32_t *)(dp + o_p);
  launch_mark_movers(d_pairs, (uint32_t)(pairs.size() / 2), (uint32_t *)c->d_mover_of_model.p, 0, c->stream);
  TouchBatchView v;
  v.n_queries = n_movers;
  v.mover_of_model = (const uint32_t *)c->d_mover_of_model.p;
  auto touches = [&](int set, int slot) {
    v.queries = (const CollisionQuery *)(dp + o_q[set]);
    v.edges = (const EdgeRec *)(dp + o_e[set]);
    v.edge_z = (const double *)(dp + o_z[set]);
    v.out = (TouchRec *)c->d_mv_out.p + (size_t)slot * n_movers;
    launch_touches(c->g, v, c->stream);
    c->stats.launches_post++;
  };
  touches(0, 0); // A: the wanted outlines against everybody where the grid has them now (what the movers still to come look like)

  // ---- move everybody (Block::UnMap + Block::Map per block, in the loop's order); every mover keeps as many
  // sequence numbers in reserve as it has blocks, for the re-maps of an undo
  std::vector<uint32_t> undo_seq(n_movers);
  for (uint32_t i = 0; i < n_movers; i++) {
    for (uint32_t k = mover_first[i]; k < mover_first[i + 1] && !rc; k++) // UnMapWithChildren
      rc = block_unmap_locked(c, blocks[k], layer);
    for (uint32_t k = mover_first[i]; k < mover_first[i + 1] && !rc; k++) // MapWithChildren
      rc = block_map_locked(c, blocks[k], c->blocks[blocks[k]].model, layer, z[2 * k], z[2 * k + 1],
                            (int)(first_pt[k + 1] - first_pt[k]), xy + 2 * (size_t)first_pt[k]);
    if (rc)
      return rc;
    undo_seq[i] = c->local_seq;
    c->local_seq += mover_first[i + 1] - mover_first[i];
  }
  if ((rc = flush_locked(c)))
    return rc;
  touches(0, 1); // B: the wanted outlines against everybody at the wanted pose (movers already through that were let go)
  touches(1, 2); // C: the present outlines against everybody at the wanted pose: who would run into a mover that is put back
  launch_mark_movers(d_pairs, (uint32_t)(pairs.size() / 2), (uint32_t *)c->d_mover_of_model.p, 1, c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(c->h_mv_out.p, c->d_mv_out.p, 3 * b_out, cudaMemcpyDeviceToHost, c->stream));
  c->stats.d2h_bytes += 3 * b_out;
  CU(c, cudaStreamSynchronize(c->stream));

  // ---- decide in order (model_position.cc:554): mover i sees the movers before it where they ended up - at the wanted
  // pose, or back at the present one - and the movers after it where the grid had them
  const TouchRec *ta = (const TouchRec *)c->h_mv_out.p, *tb = ta + n_movers, *tc = tb + n_movers;
  std::vector<uint8_t> runs_into_undone(n_movers, 0);
  bool any_hit = false;
  for (uint32_t i = 0; i < n_movers; i++) {
    if (ta[i].n_touched > kTouchMax || tb[i].n_touched > kTouchMax || tc[i].n_touched > kTouchMax)
      return fail(c, STG_RT_ENOMEM, "models_move: mover %u touches more than %u other movers; the movers are at their new poses", i, kTouchMax);
    bool hit = ground[i] || ta[i].static_hit || tb[i].static_hit || runs_into_undone[i];
    for (uint32_t k = 0; k < ta[i].n_touched && !hit; k++)
      hit = ta[i].touched[k] > i;
    for (uint32_t k = 0; k < tb[i].n_touched && !hit; k++)
      hit = tb[i].touched[k] < i && !stalled[tb[i].touched[k]];
    stalled[i] = hit ? 1 : 0;
    any_hit |= hit;
    if (hit) // i goes back to its present outline: the later movers whose wanted outline touches that one hit it
      for (uint32_t k = 0; k < tc[i].n_touched; k++)
        if (tc[i].touched[k] > i)
          runs_into_undone[tc[i].touched[k]] = 1;
  }
  if (!any_hit)
    return STG_RT_OK;
  // ---- undo the movers that hit something: UnMap + Map at the present pose, with the sequence numbers kept for them
  for (uint32_t i = 0; i < n_movers; i++) {
    if (!stalled[i])
      continue;
    c->local_seq = undo_seq[i];
    for (uint32_t k = mover_first[i]; k < mover_first[i + 1] && !rc; k++)
      rc = block_unmap_locked(c, blocks[k], layer);
    for (uint32_t k = mover_first[i]; k < mover_first[i + 1] && !rc; k++)
      rc = block_map_locked(c, blocks[k], c->blocks[blocks[k]].model, layer, z_start[2 * k], z_start[2 * k + 1],
                            (int)(first_pt_start[k + 1] - first_pt_start[k]), xy_start + 2 * (size_t)first_pt_start[k]);
    if (rc)
      return rc;
  }
  c->hold_epoch = true;
  rc = flush_locked(c);
  c->hold_epoch = false;
  return rc;
}

int stg_rt_blobs_submit(stg_rt_ctx *c, uint32_t n_finders, const stg_rt_blob_finder *finders, uint32_t n_colors,
                        const double *colors_rgb, uint32_t max_per_finder, stg_rt_blob *out, uint32_t *out_count)
{
  static_assert(sizeof(stg_rt_blob_finder) == sizeof(BlobFinderRec) && sizeof(stg_rt_blob) == sizeof(BlobRec), "blob record layouts");
  if (!c || (n_finders && (!finders || !out || !out_count)) || (n_colors && !colors_rgb) || !max_per_finder)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!n_finders)
    return STG_RT_OK;
  uint64_t beams = 0;
  for (uint32_t i = 0; i < n_finders; i++) {
    const stg_rt_blob_finder &f = finders[i];
    if (f.finder >= c->cfg.max_models || !c->model_known[f.finder] || f.layer > 1 || !f.scan_width || !f.scan_height ||
        (uint64_t)f.first_color + f.n_colors > n_colors)
      return fail(c, STG_RT_EINVAL, "blobfinder %u: unknown model %u, bad layer / image size / colours", i, f.finder);
    beams += f.scan_width;
  }
  int rc = flush_locked(c);
  if (rc || (rc = deliver_post_locked(c)))
    return rc;
  if (c->mdl_color_dirty) {
    const size_t b_col = c->mdl_color.size() * 8;
    if ((rc = dev_reserve(c, c->d_mdl_color, b_col)))
      return rc;
    CU(c, cudaMemcpyAsync(c->d_mdl_color.p, c->mdl_color.data(), b_col, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream)); // pageable source
    c->stats.h2d_bytes += b_col;
    c->mdl_color_dirty = false;
  }
  // pinned input block: finders | fans | first_beam | colours
  const size_t b_f = (size_t)n_finders * sizeof(BlobFinderRec), b_fan = (size_t)n_finders * sizeof(FanRec);
  const size_t b_first = ((size_t)(n_finders + 1) * 4 + 7) & ~(size_t)7, b_col = (size_t)n_colors * 24;
  const size_t b_in = b_f + b_fan + b_first + b_col;
  const size_t b_out = (size_t)n_finders * max_per_finder * sizeof(BlobRec), b_cnt = (size_t)n_finders * 4;
  if ((rc = pin_reserve(c, c->h_blob_in, b_in)) || (rc = dev_reserve(c, c->d_blob_in, std::max<size_t>(b_in, 8))) ||
      (rc = pin_reserve(c, c->h_blob_out, b_out)) || (rc = dev_reserve(c, c->d_blob_out, b_out)) ||
      (rc = pin_reserve(c, c->h_blob_count, b_cnt)) || (rc = dev_reserve(c, c->d_blob_count, b_cnt)) ||
      (rc = reserve_outputs(c, c->bd, beams, STG_RT_WANT_RANGES | STG_RT_WANT_HIT_MODEL)))
    return rc;
  char *hp = (char *)c->h_blob_in.p;
  memcpy(hp, finders, b_f);
  FanRec *fans = (FanRec *)(hp + b_f);
  uint32_t *first = (uint32_t *)(hp + b_f + b_fan);
  uint32_t cursor = 0;
  for (uint32_t i = 0; i < n_finders; i++) { // Raytrace(Pose(0,0,0,pan), range, fov, blob_match, NULL, false, samples), :170
    FanRec &r = fans[i];
    memset(&r, 0, sizeof(r));
    r.finder = finders[i].finder;
    r.n_samples = finders[i].scan_width;
    r.pred = STG_RT_PRED_UNRELATED;
    r.ztest = 0;
    r.layer = finders[i].layer;
    r.angles = STG_RT_ANGLES_EVEN_FOV;
    r.ox = finders[i].ox;
    r.oy = finders[i].oy;
    r.oz = finders[i].oz;
    r.oa = finders[i].oa;
    r.range = finders[i].range;
    r.fov = finders[i].fov;
    first[i] = cursor;
    cursor += finders[i].scan_width;
  }
  first[n_finders] = cursor;
  if (b_col)
    memcpy(hp + b_f + b_fan + b_first, colors_rgb, b_col);
  CU(c, cudaMemcpyAsync(c->d_blob_in.p, hp, b_in, cudaMemcpyHostToDevice, c->stream));
  c->stats.h2d_bytes += b_in;
  char *dp = (char *)c->d_blob_in.p;
  FanBatchView fv;
  memset(&fv, 0, sizeof(fv));
  fv.n_fans = n_finders;
  fv.n_beams = beams;
  fv.beam_begin = 0;
  fv.beam_end = beams;
  fv.fans = (const FanRec *)(dp + b_f);
  fv.fan_first_beam = (const uint32_t *)(dp + b_f + b_fan);
  fv.heading = (double *)c->bd.heading.p;
  fv.ranges = (double *)c->bd.ranges.p;
  fv.hit_model = (int32_t *)c->bd.hit_model.p;
  CU(c, cudaEventRecord(c->t_ev[4][0], c->stream));
  launch_fan_headings(fv, c->stream);
  launch_ray_march(c->g, fv, c->stream);
  BlobBatchView bv;
  bv.n_finders = n_finders;
  bv.max_per_finder = max_per_finder;
  bv.finders = (const BlobFinderRec *)dp;
  bv.fan_first_beam = fv.fan_first_beam;
  bv.hit_model = fv.hit_model;
  bv.ranges = fv.ranges;
  bv.mdl_color = (const double *)c->d_mdl_color.p;
  bv.colors = (const double *)(dp + b_f + b_fan + b_first);
  bv.out = (BlobRec *)c->d_blob_out.p;
  bv.out_count = (uint32_t *)c->d_blob_count.p;
  launch_blobs(c->g, bv, c->stream);
  CU(c, cudaEventRecord(c->t_ev[4][1], c->stream));
  c->t_rec[4] = true;
  CU(c, cudaGetLastError());
  c->stats.launches_post += 2;
  c->stats.launches_march += 1;
  c->stats.beams += beams;
  c->stats.fans += n_finders;
  CU(c, cudaMemcpyAsync(c->h_blob_out.p, c->d_blob_out.p, b_out, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(c->h_blob_count.p, c->d_blob_count.p, b_cnt, cudaMemcpyDeviceToHost, c->stream));
  c->stats.d2h_bytes += b_out + b_cnt;
  c->blob_job.user_out = out;
  c->blob_job.user_count = out_count;
  c->blob_job.out_bytes = b_out;
  c->blob_job.count_bytes = b_cnt;
  c->blob_job.pending = true;
  return STG_RT_OK;
}

// ---- plumbing --------------------------------------------------------------------------------------

int stg_rt_set_stream(stg_rt_ctx *c, void *cuda_stream)
{
  if (!c)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc)
    return rc;
  if ((rc = deliver_locked(c)))
    return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return STG_RT_OK;
}

void *stg_rt_get_stream(stg_rt_ctx *c) { return c ? (void *)c->stream : nullptr; }

int stg_rt_debug_trig(stg_rt_ctx *c, uint64_t n, const double *headings, double *cos_sin)
{
  if (!c || (n && (!headings || !cos_sin)))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  DevBuf dx, dcs;
  int rc = dev_reserve(c, dx, n * 8 + 8);
  if (!rc)
    rc = dev_reserve(c, dcs, n * 16 + 16);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(dx.p, headings, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
      launch_debug_trig(n, (const double *)dx.p, (double *)dcs.p, c->stream);
      e = cudaMemcpyAsync(cos_sin, dcs.p, n * 16, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
      rc = fail(c, STG_RT_ECUDA, "debug_trig: %s", cudaGetErrorString(e));
  }
  if (dx.p)
    cudaFree(dx.p);
  if (dcs.p)
    cudaFree(dcs.p);
  return rc;
}

int stg_rt_debug_headings(stg_rt_ctx *c, uint32_t n_fans, const stg_rt_fan *fans, const double *headings_in, double *headings_out)
{
  if (!c || (n_fans && (!fans || !headings_out)))
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  std::vector<uint32_t> first(n_fans + 1, 0);
  uint64_t n_in = 0;
  for (uint32_t i = 0; i < n_fans; i++) {
    if (fans[i].angles > STG_RT_ANGLES_EXPLICIT || (fans[i].angles == STG_RT_ANGLES_EXPLICIT && !headings_in))
      return fail(c, STG_RT_EINVAL, "debug_headings: fan %u: bad angle mode", i);
    first[i + 1] = first[i] + fans[i].n_samples;
    if (fans[i].angles == STG_RT_ANGLES_EXPLICIT)
      n_in = std::max<uint64_t>(n_in, (uint64_t)fans[i].first_heading + fans[i].n_samples);
  }
  const uint64_t beams = first[n_fans];
  if (!beams)
    return STG_RT_OK;
  static_assert(sizeof(stg_rt_fan) == sizeof(FanRec), "fan record");
  DevBuf dfans, dfirst, din, dout;
  int rc = dev_reserve(c, dfans, (size_t)n_fans * sizeof(FanRec));
  if (!rc) rc = dev_reserve(c, dfirst, (size_t)(n_fans + 1) * 4);
  if (!rc) rc = dev_reserve(c, din, n_in * 8 + 8);
  if (!rc) rc = dev_reserve(c, dout, beams * 8);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(dfans.p, fans, (size_t)n_fans * sizeof(FanRec), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dfirst.p, first.data(), (size_t)(n_fans + 1) * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_in)
      e = cudaMemcpyAsync(din.p, headings_in, n_in * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
      FanBatchView v;
      memset(&v, 0, sizeof(v));
      v.n_fans = n_fans;
      v.n_beams = v.beam_end = beams;
      v.fans = (const FanRec *)dfans.p;
      v.fan_first_beam = (const uint32_t *)dfirst.p;
      v.headings_in = n_in ? (const double *)din.p : nullptr;
      v.heading = (double *)dout.p;
      launch_fan_headings(v, c->stream);
      e = cudaMemcpyAsync(headings_out, dout.p, beams * 8, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
      rc = fail(c, STG_RT_ECUDA, "debug_headings: %s", cudaGetErrorString(e));
  }
  for (DevBuf *b : { &dfans, &dfirst, &din, &dout })
    if (b->p)
      cudaFree(b->p);
  return rc;
}

int stg_rt_debug_link_bandwidth(stg_rt_ctx *c, int mode, size_t bytes, int reps, double *gb_per_s)
{
  if (!c || !gb_per_s || bytes < 4096 || reps < 1 || mode < 0 || mode > 1)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = flush_locked(c);
  if (rc || (rc = deliver_locked(c)))
    return rc;
  void *host = nullptr, *dev = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMallocHost(&host, bytes);
  if (e == cudaSuccess)
    e = cudaMalloc(&dev, bytes);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(dev, 1, bytes, c->stream);
  if (e == cudaSuccess)
    e = cudaEventCreate(&e0);
  if (e == cudaSuccess)
    e = cudaEventCreate(&e1);
  float ms = 0;
  if (e == cudaSuccess) {
    memset(host, 0, bytes); // fault the pages in
    for (int r = -1; r < reps && e == cudaSuccess; r++) { // one untimed round first
      if (r == 0)
        e = cudaEventRecord(e0, c->stream);
      if (mode == 0)
        launch_host_store(host, bytes / 8, (unsigned long long)r + 2, c->stream);
      else if (e == cudaSuccess)
        e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess)
      e = cudaEventRecord(e1, c->stream);
    if (e == cudaSuccess)
      e = cudaEventSynchronize(e1);
    if (e == cudaSuccess)
      e = cudaEventElapsedTime(&ms, e0, e1);
  }
  if (e0)
    cudaEventDestroy(e0);
  if (e1)
    cudaEventDestroy(e1);
  if (dev)
    cudaFree(dev);
  if (host)
    cudaFreeHost(host);
  if (e != cudaSuccess)
    return fail(c, STG_RT_ECUDA, "debug_link_bandwidth: %s", cudaGetErrorString(e));
  *gb_per_s = (double)bytes * reps / (ms * 1e-3) / 1e9;
  return STG_RT_OK;
}

uint32_t stg_rt_ranger_rand_advance(uint32_t n_samples, uint32_t n_in_range)
{
  static bool have_spare = false; // generateGaussianNoise's static of the same name, model_ranger.cc:183
  uint32_t draws = 0;
  for (uint32_t t = 0; t < n_samples; t++, draws++)
    (void)rand(); // the heading jitter's simpleNoise(), :237
  for (uint32_t t = 0; t < n_in_range; t++) {
    (void)rand(); // range_noise * simpleNoise(), :245
    draws++;
    if (have_spare) { // generateGaussianNoise(range_noise_const), :246 -> :186-189
      have_spare = false;
    } else {
      have_spare = true;
      (void)rand(); // rand1, :193
      (void)rand(); // rand2, :197
      draws += 2;
    }
  }
  return draws;
}

int stg_rt_kernel_times(stg_rt_ctx *c, float ms[3])
{
  if (!c || !ms)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; k++) {
    ms[k] = -1.0f;
    if (c->t_rec[k])
      CU(c, cudaEventElapsedTime(&ms[k], c->t_ev[k][0], c->t_ev[k][1]));
  }
  return STG_RT_OK;
}

int stg_rt_kernel_times_ex(stg_rt_ctx *c, float *ms, int n)
{
  if (!c || !ms || n < 0)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < n; k++) {
    ms[k] = -1.0f;
    if (k < 6 && c->t_rec[k])
      CU(c, cudaEventElapsedTime(&ms[k], c->t_ev[k][0], c->t_ev[k][1]));
  }
  return STG_RT_OK;
}

int stg_rt_get_stats(stg_rt_ctx *c, stg_rt_stats *out)
{
  if (!c || !out)
    return STG_RT_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  *out = c->stats;
  return STG_RT_OK;
}

} // extern "C"
