// T1 oracle: an instrumented CPU restatement of the reference's raycast hot path.
// TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
// may load this; the product (stage_b200/, include/) never does.
//
// Pinned against the unmodified reference (T0 = oracle/_ref/libstage_ref.so via ref_harness.cc):
// tests/test_oracle_vs_reference.py replays recorded reference traces and asserts bit-identical
// range / hit model on every ray, and an identical cell-for-cell grid.
//
// What is restated (file:line under /root/reference/libstage):
//   t1_block_map      World::MapPoly            world.cc:990-1056   Cell::AddBlock    region.cc:283-290
//   t1_block_unmap    Block::UnMap              block.cc:183-189    Cell::RemoveBlock region.cc:292-299
//   t1_raytrace       World::Raytrace(Ray)      world.cc:756-963    GETCELL/GETREG/GETSREG region.hh:26-37
//   predicates        ranger_match model_ranger.cc:158-170, fiducial_raytrace_match
//                     model_fiducial.cc:103-107, blob_match model_blobfinder.cc:102-107,
//                     bumper_match model_bumper.cc:157-162, gripper_raytrace_match
//                     model_gripper.cc:329-336; Model::IsRelated model.cc:446-466 == same root
//   t1_ranger_fan     ModelRanger::Sensor::Update model_ranger.cc:211-256 (noise-free path)
//   t1_even_fan       World::Raytrace(fan)      world.cc:721-744
// On top of the reference's outputs it reports, per ray, the hit cell and the two counters that
// define the algorithmic-bytes figure: C = inner-loop cell visits
// (world.cc:840-882), R = outer-loop region probes (world.cc:818-823).
//
// Arithmetic notes: doubles throughout, no FMA contraction (build with -ffp-contract=off),
// double->int32 conversions truncate toward zero exactly where the reference's implicit
// conversions do. sin/cos come from the host libm, as in the reference (world.cc:774-775).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <set>
#include <vector>

#include "trace_format.h"

namespace {

const int RBITS = 5;                 // region.hh:13
const int REGIONWIDTH = 1 << RBITS;  // region.hh:17
const int REGIONSIZE = REGIONWIDTH * REGIONWIDTH;

struct T1Model {
  double color[4];
  uint32_t root;
  double ranger_return;
  int32_t fiducial_return, fiducial_key;
  uint32_t flags;
  bool known;
};

struct T1Block {
  uint32_t model;
  double zmin, zmax;
  std::vector<int64_t> rendered[2]; // cell keys, one per AddBlock (Block::rendered_cells)
  bool known;
};

struct T1Region {
  long count;                                  // Region::count, both layers (region.cc:19-34)
  std::vector<std::vector<uint32_t> > cells[2]; // Cell::blocks[layer], REGIONSIZE entries when allocated
  T1Region() : count(0) {}
};

inline int64_t region_key(int32_t grx, int32_t gry) { return ((int64_t)gry << 32) | (uint32_t)grx; }
inline int64_t cell_key(int32_t x, int32_t y) { return ((int64_t)y << 32) | (uint32_t)x; }

struct T1World {
  double ppm;
  std::vector<T1Model> models;
  std::vector<T1Block> blocks;
  std::unordered_map<int64_t, T1Region *> regions;

  T1Region *find_region(int32_t grx, int32_t gry) const
  {
    std::unordered_map<int64_t, T1Region *>::const_iterator it = regions.find(region_key(grx, gry));
    return it == regions.end() ? NULL : it->second;
  }
  T1Region *get_region(int32_t grx, int32_t gry)
  {
    T1Region *&r = regions[region_key(grx, gry)];
    if (!r)
      r = new T1Region;
    return r;
  }
};

inline int32_t sgn_i(int32_t a) { return a < 0 ? -1 : 1; }      // stage.hh:173
inline double sgn_d(double a) { return a < 0 ? -1.0 : 1.0; }    // stage.hh:179

// region.cc:283-290 + region.hh:77-89
void add_block_to_cell(T1World *w, uint32_t block, int layer, int32_t x, int32_t y)
{
  T1Region *reg = w->get_region(x >> RBITS, y >> RBITS);
  for (int l = 0; l < 2; l++)
    if (reg->cells[l].empty())
      reg->cells[l].resize(REGIONSIZE);
  reg->cells[layer][(x & (REGIONWIDTH - 1)) + (y & (REGIONWIDTH - 1)) * REGIONWIDTH].push_back(block);
  w->blocks[block].rendered[layer].push_back(cell_key(x, y));
  reg->count++;
}

bool predicate(const T1World *w, int kind, uint32_t hit, uint32_t finder)
{
  const T1Model &h = w->models[hit], &f = w->models[finder];
  switch (kind) {
  case TRACE_KIND_RANGER: // hit outside the finder's tree and not invisible to rangers
    return h.root != f.root && !(h.ranger_return < 0);
  case TRACE_KIND_GRIPPER:
    return hit != finder && (h.flags & TRACE_FLAG_GRIPPER_RETURN);
  default:
    return h.root != f.root;
  }
}

} // namespace

extern "C" {

typedef struct {
  double ox, oy, oz, oa, range;
  uint32_t finder;
  uint8_t kind, ztest, layer, pad;
} T1RayIn; // 48 B

typedef struct {
  double range;
  int32_t hit_model;      // -1 = none
  int32_t hit_x, hit_y;   // global cell index of the hit cell (valid when hit_model >= 0)
  uint32_t cell_visits;   // C
  uint32_t region_probes; // R
  uint32_t pad;
} T1Hit; // 32 B

void *t1_create(double ppm)
{
  T1World *w = new T1World;
  w->ppm = ppm;
  return w;
}

void t1_destroy(void *wv)
{
  T1World *w = (T1World *)wv;
  for (std::unordered_map<int64_t, T1Region *>::iterator it = w->regions.begin(); it != w->regions.end(); ++it)
    delete it->second;
  delete w;
}

void t1_model_upsert(void *wv, uint32_t id, uint32_t root, double ranger_return, int32_t fiducial_return,
                     int32_t fiducial_key, uint32_t flags)
{
  T1World *w = (T1World *)wv;
  if (id >= w->models.size()) {
    T1Model z;
    memset(&z, 0, sizeof(z));
    w->models.resize(id + 1, z);
  }
  T1Model &m = w->models[id];
  m.root = root;
  m.ranger_return = ranger_return;
  m.fiducial_return = fiducial_return;
  m.fiducial_key = fiducial_key;
  m.flags = flags;
  m.known = true;
}

// Block::Map(layer) as seen by the grid: World::MapPoly over the pixel-space polygon, then the
// z bounds (shared by both layers, block.cc:177-180).
int t1_block_map(void *wv, uint32_t block, uint32_t model, int layer, double zmin, double zmax, int n_pts,
                 const int32_t *xy)
{
  T1World *w = (T1World *)wv;
  if (layer < 0 || layer > 1 || model >= w->models.size() || !w->models[model].known)
    return -1;
  if (block >= w->blocks.size())
    w->blocks.resize(block + 1);
  T1Block &b = w->blocks[block];
  b.model = model;
  b.known = true;
  for (int i = 0; i < n_pts; i++) {
    const int32_t x0 = xy[2 * i], y0 = xy[2 * i + 1];
    const int32_t x1 = xy[2 * ((i + 1) % n_pts)], y1 = xy[2 * ((i + 1) % n_pts) + 1];
    // integer line walk, start cell included, end cell excluded (world.cc:1000-1052)
    const int32_t dx = x1 - x0, dy = y1 - y0;
    const int32_t sx = sgn_i(dx), sy = sgn_i(dy);
    const int32_t ax = std::abs(dx), ay = std::abs(dy);
    const int32_t bx = 2 * ax, by = 2 * ay;
    int32_t exy = ay - ax, n = ax + ay;
    int32_t x = x0, y = y0;
    while (n > 0) {
      add_block_to_cell(w, block, layer, x, y);
      if (exy < 0) {
        x += sx;
        exy += by;
      } else {
        y += sy;
        exy -= bx;
      }
      --n;
    }
  }
  w->blocks[block].zmin = zmin;
  w->blocks[block].zmax = zmax;
  return 0;
}

int t1_block_unmap(void *wv, uint32_t block, int layer)
{
  T1World *w = (T1World *)wv;
  if (layer < 0 || layer > 1 || block >= w->blocks.size() || !w->blocks[block].known)
    return -1;
  std::vector<int64_t> &rc = w->blocks[block].rendered[layer];
  for (size_t i = 0; i < rc.size(); i++) {
    const int32_t x = (int32_t)(uint32_t)(rc[i] & 0xffffffff), y = (int32_t)(rc[i] >> 32);
    T1Region *reg = w->find_region(x >> RBITS, y >> RBITS);
    std::vector<uint32_t> &lst =
        reg->cells[layer][(x & (REGIONWIDTH - 1)) + (y & (REGIONWIDTH - 1)) * REGIONWIDTH];
    lst.erase(std::remove(lst.begin(), lst.end(), block), lst.end()); // EraseAll
    reg->count--;
  }
  rc.clear();
  return 0;
}

// One ray; cosa/sina may be supplied (strict-mode trig check) or NULL to use libm like the reference.
static void trace_one(const T1World *w, const T1RayIn &r, const double *trig, T1Hit *out)
{
  const double ppm = w->ppm;
  out->range = r.range;
  out->hit_model = -1;
  out->hit_x = out->hit_y = 0;
  out->cell_visits = out->region_probes = 0;
  out->pad = 0;

  double gx = r.ox * ppm, gy = r.oy * ppm; // position in (fractional) cell units, world.cc:765-766
  const double x_start = gx, y_start = gy;
  const double ang = (r.oa == 0.0) ? 1e-12 : r.oa; // world.cc:773
  const double sina = trig ? trig[1] : sin(ang);
  const double cosa = trig ? trig[0] : cos(ang);
  const double tana = sina / cosa;
  const double dx = ppm * r.range * cosa, dy = ppm * r.range * sina;
  const int32_t sx = (int32_t)sgn_d(dx), sy = (int32_t)sgn_d(dy);
  const int32_t ax = (int32_t)std::abs(dx), ay = (int32_t)std::abs(dy);
  const int32_t bx = 2 * ax, by = 2 * ay;
  int32_t exy = ay - ax;
  int32_t n = ax + ay;

  // region-to-region strides of the two boundary-crossing sequences, world.cc:795-802
  const double xjumpx = sx * REGIONWIDTH, xjumpy = sx * REGIONWIDTH * tana;
  const double yjumpx = sy * REGIONWIDTH / tana, yjumpy = sy * REGIONWIDTH;
  const double xjumpdist = fabs(xjumpx) + fabs(xjumpy);
  const double yjumpdist = fabs(yjumpx) + fabs(yjumpy);

  double xcrossx = 0, xcrossy = 0, ycrossx = 0, ycrossy = 0, distX = 0, distY = 0;
  bool need_crossings = true;
  const int layer = r.layer;

  while (n > 0) {
    out->region_probes++;
    const int32_t ix0 = (int32_t)gx, iy0 = (int32_t)gy; // truncation, as GETSREG(double) does
    const int32_t grx = ix0 >> RBITS, gry = iy0 >> RBITS;
    const T1Region *reg = w->find_region(grx, gry);
    if (reg && reg->count) {
      need_crossings = true;
      int32_t cx = ix0 & (REGIONWIDTH - 1), cy = iy0 & (REGIONWIDTH - 1);
      while (cx >= 0 && cx < REGIONWIDTH && cy >= 0 && cy < REGIONWIDTH && n > 0) {
        out->cell_visits++;
        if (!reg->cells[layer].empty()) {
          const std::vector<uint32_t> &lst = reg->cells[layer][cx + cy * REGIONWIDTH];
          for (size_t k = 0; k < lst.size(); k++) {
            const T1Block &b = w->blocks[lst[k]];
            if (r.ztest && (r.oz < b.zmin || r.oz > b.zmax))
              continue;
            if (predicate(w, r.kind, b.model, r.finder)) {
              out->hit_model = (int32_t)b.model;
              out->hit_x = grx * REGIONWIDTH + cx;
              out->hit_y = gry * REGIONWIDTH + cy;
              if (ax > ay)
                out->range = fabs((gx - x_start) / cosa) / ppm;
              else
                out->range = fabs((gy - y_start) / sina) / ppm;
              return;
            }
          }
        }
        if (exy < 0) {
          gx += sx;
          exy += by;
          cx += sx;
        } else {
          gy += sy;
          exy -= bx;
          cy += sy;
        }
        --n;
      }
    } else {
      if (need_crossings) {
        need_crossings = false;
        const int32_t ix = (int32_t)gx, iy = (int32_t)gy;
        double regionx = ix / REGIONWIDTH * REGIONWIDTH;
        double regiony = iy / REGIONWIDTH * REGIONWIDTH;
        if ((gx < 0) && (ix % REGIONWIDTH))
          regionx -= REGIONWIDTH;
        if ((gy < 0) && (iy % REGIONWIDTH))
          regiony -= REGIONWIDTH;
        const double xdx = sx < 0 ? regionx - gx - 1.0 : regionx + REGIONWIDTH - gx;
        const double xdy = xdx * tana;
        const double ydy = sy < 0 ? regiony - gy - 1.0 : regiony + REGIONWIDTH - gy;
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
        n -= distX; // int32 -= double: converted through double, truncated (world.cc:931)
        xcrossx += xjumpx;
        xcrossy += xjumpy;
        distY -= distX;
        distX = xjumpdist;
      } else {
        gx = ycrossx;
        gy = ycrossy;
        n -= distY;
        ycrossx += yjumpx;
        ycrossy += yjumpy;
        distX -= distY;
        distY = yjumpdist;
      }
    }
  }
}

// (cos, sin) of each heading exactly as World::Raytrace would compute them on this host
// (world.cc:773-775: a heading of exactly 0 is replaced by 1e-12). out = 2 doubles per angle.
void t1_libm_trig(uint64_t n, const double *headings, double *out)
{
  for (uint64_t i = 0; i < n; i++) {
    const double ang = (headings[i] == 0.0) ? 1e-12 : headings[i];
    out[2 * i] = cos(ang);
    out[2 * i + 1] = sin(ang);
  }
}

// trig: NULL, or 2 doubles per ray (cos, sin) to use instead of libm.
void t1_raytrace(void *wv, uint64_t n, const T1RayIn *in, const double *trig, T1Hit *out)
{
  const T1World *w = (const T1World *)wv;
  for (uint64_t i = 0; i < n; i++)
    trace_one(w, in[i], trig ? trig + 2 * i : NULL, out + i);
}

// ModelRanger::Sensor::Update, noise-free (model_ranger.cc:211-256). origin = the global pose of
// the first ray (rayorg after LocalToGlobal, :226-230). Writes sample_count entries of each output.
// The float round-trip of the running angle (:237-241,254) is part of the algorithm.
void t1_ranger_fan(void *wv, uint32_t finder, int layer, const double *origin_xyza, double range_max,
                   double fov, uint32_t sample_count, double *ranges, double *intensities,
                   double *bearings, T1Hit *hits)
{
  const T1World *w = (const T1World *)wv;
  const double sample_incr = fov / std::max(sample_count - 1, (unsigned int)1);
  const double start_angle = (sample_count > 1 ? -fov / 2.0 : 0.0);
  T1RayIn r;
  memset(&r, 0, sizeof(r));
  r.ox = origin_xyza[0];
  r.oy = origin_xyza[1];
  r.oz = origin_xyza[2];
  r.oa = origin_xyza[3];
  r.range = range_max;
  r.finder = finder;
  r.kind = TRACE_KIND_RANGER;
  r.ztest = 1;
  r.layer = (uint8_t)layer;
  for (uint32_t t = 0; t < sample_count; t++) {
    const float saved = (float)r.oa;
    const float distorted = (float)(r.oa + sample_incr * 0.0 * 0.0 * 0.5); // angle_noise == 0
    r.oa = distorted;
    T1Hit h;
    trace_one(w, r, NULL, &h);
    r.oa = saved;
    ranges[t] = h.range;
    intensities[t] = h.hit_model >= 0 ? w->models[h.hit_model].ranger_return : 0.0;
    bearings[t] = start_angle + ((double)t) * sample_incr;
    if (hits)
      hits[t] = h;
    r.oa += sample_incr;
  }
}

// World::Raytrace(gpose, range, fov, ...) (world.cc:721-744): evenly spaced fan, double angles.
void t1_even_fan(void *wv, uint32_t finder, int layer, int kind, int ztest, const double *gpose_xyza,
                 double range, double fov, uint32_t sample_count, T1Hit *hits)
{
  const T1World *w = (const T1World *)wv;
  const double starta = fov / 2.0 - gpose_xyza[3];
  T1RayIn r;
  memset(&r, 0, sizeof(r));
  r.ox = gpose_xyza[0];
  r.oy = gpose_xyza[1];
  r.oz = gpose_xyza[2];
  r.range = range;
  r.finder = finder;
  r.kind = (uint8_t)kind;
  r.ztest = (uint8_t)ztest;
  r.layer = (uint8_t)layer;
  for (uint32_t s = 0; s < sample_count; s++) {
    r.oa = (s * fov / (double)(sample_count - 1)) - starta;
    trace_one(w, r, NULL, hits + s);
  }
}

// Model::color, read by the blobfinder through RaytraceResult::color (world.cc:853).
void t1_model_color(void *wv, uint32_t id, const double *rgba)
{
  T1World *w = (T1World *)wv;
  if (id < w->models.size())
    memcpy(w->models[id].color, rgba, sizeof(double) * 4);
}

static double normalize_angle(double a) // stage.hh:163-170
{
  while (a < -M_PI)
    a += 2.0 * M_PI;
  while (a > M_PI)
    a -= 2.0 * M_PI;
  return a;
}

typedef struct {
  uint32_t model;
  int32_t fiducial_return, fiducial_key;
  uint32_t reserved;
  double x, y, z, a;
  double sx, sy, sz;
} T1FidTarget; // 72 B, same layout as stg_rt_fid_target

typedef struct {
  uint32_t finder;
  int32_t fiducial_key;
  uint8_t ignore_zloc, layer, pad[6];
  double x, y, z, a;
  double ox, oy, oz, oa;
  double max_range_anon, max_range_id, fov;
} T1FidFinder; // 104 B, same layout as stg_rt_fid_finder

typedef struct {
  uint32_t target;
  int32_t id;
  double range, bearing;
  double geom[4], pose[4];
} T1Fiducial; // 88 B, same layout as stg_rt_fiducial

// ModelFiducial::Update + AddModelIfVisible (model_fiducial.cc:109-293) for one finder. targets are
// World::models_with_fiducials in ascending Model* order. Returns the number of fiducials found
// (records beyond max_out are counted but not stored).
uint32_t t1_fiducial_update(void *wv, uint32_t n_targets, const T1FidTarget *targets, const T1FidFinder *fd,
                            uint32_t max_out, T1Fiducial *out)
{
  const T1World *w = (const T1World *)wv;
  uint32_t n = 0;
  const double rng = fd->max_range_anon;
  for (uint32_t t = 0; t < n_targets; t++) {
    const T1FidTarget &him = targets[t];
    // the sorted-set window query (:245-283): within +-rng on both axes, bounds inclusive
    if (him.x < fd->x - rng || him.x > fd->x + rng || him.y < fd->y - rng || him.y > fd->y + rng)
      continue;
    if (fd->fiducial_key != him.fiducial_key) // :117
      continue;
    const double dx = him.x - fd->x, dy = him.y - fd->y;
    const double range = hypot(dy, dx);
    if (range >= fd->max_range_anon) // :134
      continue;
    const double bearing = atan2(dy, dx);
    const double dtheta = normalize_angle(bearing - fd->a);
    if (fabs(dtheta) > fd->fov / 2.0) // :147
      continue;
    if (w->models[him.model].root == w->models[fd->finder].root) // IsRelated, :152
      continue;
    T1RayIn r;
    memset(&r, 0, sizeof(r));
    r.ox = fd->ox; // LocalToGlobal(Pose(0,0,0,dtheta)), stage.hh:2296
    r.oy = fd->oy;
    r.oz = fd->oz;
    r.oa = normalize_angle(fd->oa + dtheta);
    r.range = fd->max_range_anon;
    r.finder = fd->finder;
    r.kind = TRACE_KIND_UNRELATED;
    r.ztest = 1;
    r.layer = fd->layer;
    T1Hit h;
    trace_one(w, r, NULL, &h);
    int32_t seen = h.hit_model;
    if (fd->ignore_zloc && seen < 0) // :184
      seen = (int32_t)him.model;
    if (seen != (int32_t)him.model) // :192
      continue;
    if (n < max_out) {
      T1Fiducial &f = out[n];
      f.target = t;
      f.id = range < fd->max_range_id ? him.fiducial_return : 0; // :219
      f.range = range;
      f.bearing = dtheta;
      f.geom[0] = him.sx;
      f.geom[1] = him.sy;
      f.geom[2] = him.sz;
      f.geom[3] = normalize_angle(him.a - fd->a);
      f.pose[0] = him.x;
      f.pose[1] = him.y;
      f.pose[2] = him.z;
      f.pose[3] = him.a;
    }
    n++;
  }
  return n;
}

typedef struct {
  uint32_t model, left, top, right, bottom, pad;
  double range;
} T1Blob; // 32 B, same layout as stg_rt_blob

static bool color_match(const double *a, const double *b) // ColorMatchIgnoreAlpha, model_blobfinder.cc:109-113
{
  const double epsilon = 1e-5;
  return fabs(a[0] - b[0]) < epsilon && fabs(a[1] - b[1]) < epsilon && fabs(a[2] - b[2]) < epsilon;
}

// ModelBlobfinder::Update (model_blobfinder.cc:165-250). origin = LocalToGlobal(Pose(0,0,0,pan)).
uint32_t t1_blob_update(void *wv, uint32_t finder, int layer, const double *origin_xyza, double range, double fov,
                        uint32_t scan_width, uint32_t scan_height, uint32_t n_colors, const double *colors_rgb,
                        uint32_t max_out, T1Blob *out)
{
  const T1World *w = (const T1World *)wv;
  std::vector<T1Hit> samples(scan_width);
  t1_even_fan(wv, finder, layer, TRACE_KIND_UNRELATED, 0, origin_xyza, range, fov, scan_width, samples.data());
  const double yRadsPerPixel = fov / scan_height;
  uint32_t n = 0;
  for (unsigned int s = 0; s < scan_width; s++) {
    if (samples[s].hit_model < 0)
      continue;
    unsigned int right = s;
    const int32_t blob_model = samples[s].hit_model;
    const double *blobcol = w->models[blob_model].color;
    while (s < scan_width && samples[s].hit_model >= 0 && color_match(w->models[samples[s].hit_model].color, blobcol))
      s++;
    unsigned int left = s - 1;
    if (n_colors) {
      bool found = false;
      for (unsigned int c = 0; c < n_colors; c++)
        if (color_match(blobcol, colors_rgb + 3 * c)) {
          found = true;
          break;
        }
      if (!found)
        continue;
    }
    double robotHeight = 0.6;
    double brange = 0;
    for (unsigned int t = right; t <= left; t++)
      brange += samples[t].range;
    brange /= left - right + 1;
    double startyangle = atan2(robotHeight / 2.0, brange);
    double endyangle = -startyangle;
    int blobtop = scan_height / 2 - (int)(startyangle / yRadsPerPixel);
    int blobbottom = scan_height / 2 - (int)(endyangle / yRadsPerPixel);
    blobtop = std::max(blobtop, 0);
    blobbottom = std::min(blobbottom, (int)scan_height);
    if (n < max_out) {
      T1Blob &b = out[n];
      b.model = (uint32_t)blob_model;
      b.left = scan_width - left - 1;
      b.top = (uint32_t)blobtop;
      b.right = scan_width - right - 1;
      b.bottom = (uint32_t)blobbottom;
      b.pad = 0;
      b.range = brange;
    }
    n++;
  }
  return n;
}

// Same record format as ref_dump_grid (oracle/ref_harness.cc): {layer, x, y, pos, block} int32[5],
// sorted (layer, y, x, pos) by the caller. Returns the number of records.
int64_t t1_dump_grid(void *wv, int32_t *out, int64_t cap, int64_t *counts, int64_t counts_cap, int64_t *n_counts)
{
  const T1World *w = (const T1World *)wv;
  int64_t n = 0, nc = 0;
  for (std::unordered_map<int64_t, T1Region *>::const_iterator it = w->regions.begin(); it != w->regions.end(); ++it) {
    const int32_t grx = (int32_t)(uint32_t)(it->first & 0xffffffff), gry = (int32_t)(it->first >> 32);
    const T1Region *reg = it->second;
    if (reg->count) {
      if (counts && nc < counts_cap) {
        counts[3 * nc] = grx;
        counts[3 * nc + 1] = gry;
        counts[3 * nc + 2] = reg->count;
      }
      nc++;
    }
    for (int layer = 0; layer < 2; layer++) {
      if (reg->cells[layer].empty())
        continue;
      for (int c = 0; c < REGIONSIZE; c++) {
        const std::vector<uint32_t> &lst = reg->cells[layer][c];
        for (size_t p = 0; p < lst.size(); p++) {
          if (out && n < cap) {
            int32_t *o = out + 5 * n;
            o[0] = layer;
            o[1] = grx * REGIONWIDTH + (c & (REGIONWIDTH - 1));
            o[2] = gry * REGIONWIDTH + (c >> RBITS);
            o[3] = (int32_t)p;
            o[4] = (int32_t)lst[p];
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

// n times { Block::UnMap(layer); Block::Map(layer) } in one call (what ModelPosition::Move does to the grid for
// n moving models, model_position.cc:549-552): the loop a caller of t1_block_unmap / t1_block_map would write.
int t1_blocks_remap(void *wv, uint32_t n, const uint32_t *blocks, const uint32_t *models, int layer, const double *z,
                    const uint32_t *first_pt, const int32_t *xy)
{
  for (uint32_t i = 0; i < n; i++) {
    const T1World *w = (const T1World *)wv;
    if (blocks[i] < w->blocks.size() && w->blocks[blocks[i]].known)
      t1_block_unmap(wv, blocks[i], layer);
    if (t1_block_map(wv, blocks[i], models[i], layer, z[2 * i], z[2 * i + 1], (int)(first_pt[i + 1] - first_pt[i]), xy + 2 * (size_t)first_pt[i]))
      return -1;
  }
  return 0;
}

// An order-independent digest of the whole grid, list order included: the sum over every (cell, block) entry of a
// 64-bit mix of (layer, x, y, position in Cell::blocks[layer], block). Two grids agree on it iff (up to 2^-64) they hold
// the same blocks in the same order in every cell. The device computes the same sum over its own tables
// (stg_rt_grid_hash), so replicas can be compared with each other and with this without moving the grid.
static inline uint64_t t1_mix64(uint64_t z)
{
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}

void t1_grid_hash(void *wv, uint64_t out[5])
{
  const T1World *w = (const T1World *)wv;
  for (int i = 0; i < 5; i++)
    out[i] = 0;
  for (std::unordered_map<int64_t, T1Region *>::const_iterator it = w->regions.begin(); it != w->regions.end(); ++it) {
    const int32_t grx = (int32_t)(uint32_t)(it->first & 0xffffffff), gry = (int32_t)(it->first >> 32);
    const T1Region *reg = it->second;
    if (reg->count)
      out[4] += t1_mix64(t1_mix64(((uint64_t)(uint32_t)grx << 32) | (uint32_t)gry) ^ ((uint64_t)reg->count * 0x9e3779b97f4a7c15ull));
    for (int layer = 0; layer < 2; layer++) {
      if (reg->cells[layer].empty())
        continue;
      for (int c = 0; c < REGIONSIZE; c++) {
        const std::vector<uint32_t> &lst = reg->cells[layer][c];
        const int32_t x = grx * REGIONWIDTH + (c & (REGIONWIDTH - 1)), y = gry * REGIONWIDTH + (c >> RBITS);
        uint32_t pos = 0;
        for (size_t p = 0; p < lst.size(); p++) {
          bool seen = false; // a block rendered twice into a cell (Map on a mapped block) counts where it first stands
          for (size_t q = 0; q < p && !seen; q++)
            seen = lst[q] == lst[p];
          if (seen)
            continue;
          const uint64_t a = ((uint64_t)(uint32_t)x << 32) | (uint32_t)y, b = ((uint64_t)lst[p] << 32) | pos;
          out[layer] += t1_mix64(t1_mix64(a) ^ (b * 0x9e3779b97f4a7c15ull));
          out[2 + layer]++;
          pos++;
        }
      }
    }
  }
}

typedef struct {
  uint64_t rays, range_mismatch, model_mismatch, maps, unmaps, ticks;
  uint64_t cell_visits, region_probes;
  uint64_t scans, scan_mismatch;
  uint64_t fiducial_updates, fiducial_mismatch, blob_updates, blob_mismatch;
  double max_range_err;
  uint64_t collision_tests, collision_mismatch, collision_hits;
  uint64_t toucher_calls, toucher_mismatch, toucher_models;
} T1ReplayStats;

// Model::TestCollision (model.cc:666-679) over BlockGroup::TestCollision (blockgroup.cc:41-50) and
// Block::TestCollision (block.cc:129-168): `blocks` are the dense ids of the model's own blocks and
// then its descendants', depth first - the order those three visit them in. Returns the id of the
// first model found in the way, or -1.
//   per block: nothing to test unless its own model has obstacle_return (:136); a block reaching
//   below the floor collides with the ground model, id 0 (:137-138, World::GetGround); otherwise
//   every cell the block is rendered into on `layer`, in rendering order (Block::rendered_cells,
//   filled by MapPoly edge by edge), and every block listed in that cell, in list order: the first
//   whose model is another one, is an obstacle, is not related (same tree) and overlaps in z (:153-157).
int32_t t1_test_collision(void *wv, int layer, uint32_t n_blocks, const uint32_t *blocks)
{
  const T1World *w = (const T1World *)wv;
  for (uint32_t i = 0; i < n_blocks; i++) {
    if (blocks[i] >= w->blocks.size() || !w->blocks[blocks[i]].known)
      continue;
    const T1Block &b = w->blocks[blocks[i]];
    const T1Model &own = w->models[b.model];
    if (!(own.flags & TRACE_FLAG_OBSTACLE_RETURN))
      continue;
    if (b.zmin < 0)
      return 0;
    const std::vector<int64_t> &cells = b.rendered[layer];
    for (size_t c = 0; c < cells.size(); c++) {
      const int32_t x = (int32_t)(uint32_t)cells[c], y = (int32_t)(cells[c] >> 32); // cell_key
      const T1Region *reg = w->find_region(x >> RBITS, y >> RBITS);
      if (!reg || reg->cells[layer].empty())
        continue;
      const std::vector<uint32_t> &list = reg->cells[layer][(x & (REGIONWIDTH - 1)) + (y & (REGIONWIDTH - 1)) * REGIONWIDTH];
      for (size_t k = 0; k < list.size(); k++) {
        const T1Block &tb = w->blocks[list[k]];
        const T1Model &tm = w->models[tb.model];
        if (tb.model != b.model && (tm.flags & TRACE_FLAG_OBSTACLE_RETURN) && tm.root != own.root && tb.zmin <= b.zmax &&
            tb.zmax >= b.zmin)
          return (int32_t)tb.model;
      }
    }
  }
  return -1;
}

// Model::AppendTouchingModels (model.cc:661-664) -> BlockGroup::AppendTouchingModels (blockgroup.cc:35-39) ->
// Block::AppendTouchingModels (block.cc:116-127): `blocks` are the dense ids of the model's own blocks (its
// children are not visited). For every cell a block is rendered into on `layer` (= World::updates % 2) and
// every block listed in that cell: its model joins the set unless it is related to this one (same tree -
// which covers the block itself). No obstacle_return, no z test. The reference collects a std::set<Model*>;
// the ids come back ascending, at most `cap` of them; returns how many there are.
uint32_t t1_touching_models(void *wv, int layer, uint32_t n_blocks, const uint32_t *blocks, uint32_t cap, uint32_t *out)
{
  const T1World *w = (const T1World *)wv;
  std::set<uint32_t> found;
  for (uint32_t i = 0; i < n_blocks; i++) {
    if (blocks[i] >= w->blocks.size() || !w->blocks[blocks[i]].known)
      continue;
    const T1Block &b = w->blocks[blocks[i]];
    const T1Model &own = w->models[b.model];
    const std::vector<int64_t> &cells = b.rendered[layer];
    for (size_t c = 0; c < cells.size(); c++) {
      const int32_t x = (int32_t)(uint32_t)cells[c], y = (int32_t)(cells[c] >> 32);
      const T1Region *reg = w->find_region(x >> RBITS, y >> RBITS);
      if (!reg || reg->cells[layer].empty())
        continue;
      const std::vector<uint32_t> &list = reg->cells[layer][(x & (REGIONWIDTH - 1)) + (y & (REGIONWIDTH - 1)) * REGIONWIDTH];
      for (size_t k = 0; k < list.size(); k++) {
        const T1Block &tb = w->blocks[list[k]];
        if (w->models[tb.model].root != own.root) // !IsRelated, model.cc:446-466
          found.insert(tb.model);
      }
    }
  }
  uint32_t n = 0;
  for (std::set<uint32_t>::const_iterator it = found.begin(); it != found.end(); ++it, n++)
    if (out && n < cap)
      out[n] = *it;
  return n;
}

// Replay a recorded reference trace through T1 and compare every ray with the reference's answer.
// Returns the world (caller may dump its grid) or NULL on a malformed file.
void *t1_replay(const char *path, T1ReplayStats *st)
{
  memset(st, 0, sizeof(*st));
  FILE *f = fopen(path, "rb");
  if (!f)
    return NULL;
  TraceHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, STG_TRACE_MAGIC, 8)) {
    fclose(f);
    return NULL;
  }
  std::vector<TraceModel> models(h.n_models);
  std::vector<TraceEvent> events(h.n_events);
  std::vector<int32_t> verts(2 * (size_t)h.n_verts);
  std::vector<TraceRay> rays(h.n_rays);
  std::vector<double> aux(h.n_aux);
  bool ok = fread(models.data(), sizeof(TraceModel), models.size(), f) == models.size() &&
            fread(events.data(), sizeof(TraceEvent), events.size(), f) == events.size() &&
            fread(verts.data(), sizeof(int32_t), verts.size(), f) == verts.size() &&
            fread(rays.data(), sizeof(TraceRay), rays.size(), f) == rays.size() &&
            fread(aux.data(), sizeof(double), aux.size(), f) == aux.size();
  fclose(f);
  if (!ok)
    return NULL;
  void *w = t1_create(h.ppm);
  for (size_t i = 0; i < models.size(); i++) {
    t1_model_upsert(w, models[i].id, models[i].root, models[i].ranger_return, models[i].fiducial_return,
                    models[i].fiducial_key, models[i].flags);
    t1_model_color(w, models[i].id, models[i].color);
  }
  for (size_t i = 0; i < events.size(); i++) {
    const TraceEvent &e = events[i];
    if (e.type == TRACE_EV_MAP) {
      t1_block_map(w, e.block, e.model, e.layer, e.zmin, e.zmax, e.n_pts, &verts[2 * (size_t)e.first]);
      st->maps++;
    } else if (e.type == TRACE_EV_UNMAP) {
      t1_block_unmap(w, e.block, e.layer);
      st->unmaps++;
    } else if (e.type == TRACE_EV_TICK) {
      st->ticks++;
    } else if (e.type == TRACE_EV_COLLISION) {
      std::vector<uint32_t> ids(e.n_pts);
      for (uint32_t k = 0; k < e.n_pts; k++)
        ids[k] = (uint32_t)aux[e.first + k];
      st->collision_tests++;
      if (t1_test_collision(w, e.layer, e.n_pts, ids.data()) != (int32_t)e.block - 1)
        st->collision_mismatch++;
      if (e.block)
        st->collision_hits++;
    } else if (e.type == TRACE_EV_TOUCHERS) {
      std::vector<uint32_t> ids(e.n_pts), got(e.block + 1);
      for (uint32_t k = 0; k < e.n_pts; k++)
        ids[k] = (uint32_t)aux[e.first + k];
      const uint32_t n = t1_touching_models(w, e.layer, e.n_pts, ids.data(), e.block + 1, got.data());
      bool same = n == e.block;
      for (uint32_t k = 0; same && k < n; k++)
        same = got[k] == (uint32_t)aux[e.first + e.n_pts + k];
      st->toucher_calls++;
      st->toucher_models += e.block;
      if (!same)
        st->toucher_mismatch++;
    } else if (e.type == TRACE_EV_FIDUCIAL) {
      const double *a = &aux[e.first];
      const uint32_t nt = (uint32_t)a[13], nf = e.block;
      T1FidFinder fd;
      memset(&fd, 0, sizeof(fd));
      fd.finder = e.model;
      fd.layer = e.layer;
      fd.x = a[0]; fd.y = a[1]; fd.z = a[2]; fd.a = a[3];
      fd.ox = a[4]; fd.oy = a[5]; fd.oz = a[6]; fd.oa = a[7];
      fd.max_range_anon = a[8]; fd.max_range_id = a[9]; fd.fov = a[10];
      fd.fiducial_key = (int32_t)a[11];
      fd.ignore_zloc = a[12] != 0;
      std::vector<T1FidTarget> tg(nt);
      for (uint32_t k = 0; k < nt; k++) {
        const double *t = a + TRACE_FID_HEADER_DOUBLES + TRACE_FID_TARGET_DOUBLES * k;
        memset(&tg[k], 0, sizeof(T1FidTarget));
        tg[k].model = (uint32_t)t[0];
        tg[k].fiducial_return = (int32_t)t[1];
        tg[k].fiducial_key = (int32_t)t[2];
        tg[k].x = t[3]; tg[k].y = t[4]; tg[k].z = t[5]; tg[k].a = t[6];
        tg[k].sx = t[7]; tg[k].sy = t[8]; tg[k].sz = t[9];
      }
      std::vector<T1Fiducial> found(nt + 1);
      const uint32_t n = t1_fiducial_update(w, nt, tg.data(), &fd, nt + 1, found.data());
      st->fiducial_updates++;
      bool same = n == nf;
      const double *ref = a + TRACE_FID_HEADER_DOUBLES + TRACE_FID_TARGET_DOUBLES * nt;
      for (uint32_t k = 0; same && k < n; k++) {
        const double *r = ref + TRACE_FID_RESULT_DOUBLES * k;
        const T1Fiducial &f = found[k];
        same = tg[f.target].model == (uint32_t)r[0] && f.id == (int32_t)r[1] && f.range == r[2] && f.bearing == r[3] &&
               !memcmp(f.geom, r + 4, 32) && !memcmp(f.pose, r + 8, 32);
      }
      if (!same)
        st->fiducial_mismatch++;
    } else if (e.type == TRACE_EV_BLOB) {
      const double *a = &aux[e.first];
      const uint32_t nc = (uint32_t)a[8], nb = e.block, sw = (uint32_t)a[6], sh = (uint32_t)a[7];
      std::vector<T1Blob> blobs(sw + 1);
      const uint32_t n = t1_blob_update(w, e.model, e.layer, a, a[4], a[5], sw, sh, nc, a + TRACE_BLOB_HEADER_DOUBLES, sw + 1,
                                        blobs.data());
      st->blob_updates++;
      bool same = n == nb;
      const double *ref = a + TRACE_BLOB_HEADER_DOUBLES + 3 * nc;
      for (uint32_t k = 0; same && k < n; k++) {
        const double *r = ref + TRACE_BLOB_RESULT_DOUBLES * k;
        const T1Blob &b = blobs[k];
        const double *col = ((const T1World *)w)->models[b.model].color;
        same = col[0] == r[0] && col[1] == r[1] && col[2] == r[2] && b.left == (uint32_t)r[3] && b.top == (uint32_t)r[4] &&
               b.right == (uint32_t)r[5] && b.bottom == (uint32_t)r[6] && b.range == r[7];
      }
      if (!same)
        st->blob_mismatch++;
    } else if (e.type == TRACE_EV_RANGER_SCAN) {
      const uint32_t n = e.n_pts;
      const double *a = &aux[e.first];
      std::vector<double> rg(n), in(n), be(n);
      t1_ranger_fan(w, e.model, e.layer, a, a[4], a[5], n, rg.data(), in.data(), be.data(), NULL);
      const double *ref = a + TRACE_SCAN_HEADER_DOUBLES;
      st->scans++;
      if (memcmp(rg.data(), ref, 8 * n) || memcmp(in.data(), ref + n, 8 * n) || memcmp(be.data(), ref + 2 * n, 8 * n))
        st->scan_mismatch++;
    } else if (e.type == TRACE_EV_RAYS) {
      for (uint32_t k = 0; k < e.model; k++) {
        const TraceRay &tr = rays[e.first + k];
        T1RayIn in;
        memset(&in, 0, sizeof(in));
        in.ox = tr.ox;
        in.oy = tr.oy;
        in.oz = tr.oz;
        in.oa = tr.oa;
        in.range = tr.range;
        in.finder = tr.finder;
        in.kind = tr.kind;
        in.ztest = tr.ztest;
        in.layer = tr.layer;
        T1Hit hit;
        trace_one((const T1World *)w, in, NULL, &hit);
        st->rays++;
        st->cell_visits += hit.cell_visits;
        st->region_probes += hit.region_probes;
        if (memcmp(&hit.range, &tr.res_range, 8) != 0) {
          st->range_mismatch++;
          st->max_range_err = std::max(st->max_range_err, fabs(hit.range - tr.res_range));
        }
        if (hit.hit_model != tr.hit_model)
          st->model_mismatch++;
      }
    }
  }
  return w;
}

} // extern "C"
