// Host check of stage_b200/csrc/rt_walk.h: the closed forms of the region walk against the step-by-step loop the
// reference runs (world.cc:823-882 in run form). Built and run by tests/test_walk_closed_form.py. Prints "ok <cases>".
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../stage_b200/csrc/rt_walk.h"

using namespace stg::walk;

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd()
{
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 32);
}

static long cases = 0;

// one start state: the loop's exit, and every state on the way as a jump target
static int check(int32_t ap0, int32_t cp0, int32_t E0, int32_t b_run, int32_t b_cross)
{
  int32_t ap = ap0, cp = cp0, E = E0;
  int32_t first_run[kWidth + 1], first_cross[kWidth + 1]; // step index at which a coordinate is first reached
  int32_t st_ap[2 * kWidth + 2], st_cp[2 * kWidth + 2], st_E[2 * kWidth + 2];
  for (int i = 0; i <= kWidth; i++)
    first_run[i] = first_cross[i] = -1;
  int k = 0;
  st_ap[0] = ap; st_cp[0] = cp; st_E[0] = E;
  while (ap < kWidth && cp < kWidth) {
    if (E < 0) {
      ap += 1;
      E += b_run;
      k++;
      first_run[ap] = k;
    } else {
      cp += 1;
      E -= b_cross;
      k++;
      first_cross[cp] = k;
    }
    st_ap[k] = ap; st_cp[k] = cp; st_E[k] = E;
  }
  int32_t xa, xc, xe;
  region_exit(ap0, cp0, E0, b_run, b_cross, xa, xc, xe);
  cases++;
  if (xa != ap || xc != cp || xe != E) {
    printf("region_exit(%d,%d,%d,%d,%d) = (%d,%d,%d), loop (%d,%d,%d)\n", ap0, cp0, E0, b_run, b_cross, xa, xc, xe, ap, cp, E);
    return 1;
  }
  for (int32_t c = cp0 + 1; c < kWidth; c++) {
    if (first_cross[c] < 0)
      break;
    const int s = first_cross[c];
    jump_cross(ap0, cp0, E0, b_run, b_cross, c - cp0, xa, xc, xe);
    cases++;
    if (xa != st_ap[s] || xc != st_cp[s] || xe != st_E[s]) {
      printf("jump_cross(%d,%d,%d,%d,%d; %d) = (%d,%d,%d), loop (%d,%d,%d)\n", ap0, cp0, E0, b_run, b_cross, c - cp0, xa, xc, xe, st_ap[s],
             st_cp[s], st_E[s]);
      return 1;
    }
  }
  for (int32_t a = ap0 + 1; a < kWidth; a++) {
    if (first_run[a] < 0)
      break;
    const int s = first_run[a];
    jump_run(ap0, cp0, E0, b_run, b_cross, a - ap0, xa, xc, xe);
    cases++;
    if (xa != st_ap[s] || xc != st_cp[s] || xe != st_E[s]) {
      printf("jump_run(%d,%d,%d,%d,%d; %d) = (%d,%d,%d), loop (%d,%d,%d)\n", ap0, cp0, E0, b_run, b_cross, a - ap0, xa, xc, xe, st_ap[s],
             st_cp[s], st_E[s]);
      return 1;
    }
  }
  return 0;
}

int main()
{
  // every slope of a short ray, every start cell on a coarse lattice and on the borders, every error term the walk can hold
  for (int32_t major = 1; major <= 40; major++)
    for (int32_t minor = 0; minor <= major; minor++) {
      const int32_t b_run = 2 * minor, b_cross = 2 * major;
      for (int32_t E0 = -b_cross; E0 < (b_run > 0 ? b_run : 1); E0++)
        for (int32_t ap0 = 0; ap0 < kWidth; ap0 += (ap0 < 2 || ap0 >= 29) ? 1 : 3)
          for (int32_t cp0 = 0; cp0 < kWidth; cp0 += (cp0 < 2 || cp0 >= 29) ? 1 : 3)
            if (check(ap0, cp0, E0, b_run, b_cross))
              return 1;
    }
  // random rays up to the longest the closed forms serve
  for (long t = 0; t < 3000000; t++) {
    const int shift = (int)(rnd() % 24);
    int32_t major = 1 + (int32_t)(rnd() % (uint32_t)kMaxMajorCells >> shift);
    if (major > kMaxMajorCells)
      major = kMaxMajorCells;
    int32_t minor;
    switch (rnd() % 4) {
    case 0: minor = major - (int32_t)(rnd() % 3 % (uint32_t)(major + 1)); break; // near the diagonal
    case 1: minor = (int32_t)(rnd() % 3); break;                                  // near an axis
    default: minor = (int32_t)(rnd() % (uint32_t)(major + 1)); break;
    }
    if (minor > major)
      minor = major;
    if (minor < 0)
      minor = 0;
    const int32_t b_run = 2 * minor, b_cross = 2 * major;
    const int64_t span = (int64_t)b_cross + (b_run > 0 ? b_run : 1);
    const int32_t E0 = (int32_t)(-(int64_t)b_cross + (int64_t)(((uint64_t)rnd() << 32 | rnd()) % (uint64_t)span));
    if (check((int32_t)(rnd() % kWidth), (int32_t)(rnd() % kWidth), E0, b_run, b_cross))
      return 1;
  }
  printf("ok %ld\n", cases);
  return 0;
}
