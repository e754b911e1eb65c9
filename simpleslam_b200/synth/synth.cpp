// Synthetic lidar world for benchmarks and tests (bench/test utility — not on the registration path).
// Deterministic procedural scene: per 200 m tile a ground plane z = 0, ~40 axis-aligned
// boxes (footprint 5-30 m, height 3-20 m) and ~200 vertical cylinders (r = 0.3 m, h = 8 m). Analytic ray
// casting, counter-based RNG (splitmix64 of (seed, ray id)), Gaussian range noise. Points are written as
// pcl::PointXYZI-compatible 32-byte records (x y z 1 | intensity 0 0 0).
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <vector>
#include <algorithm>
#include <omp.h>

namespace {
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double u01(uint64_t seed, uint64_t a, uint64_t b) {
  uint64_t h = splitmix64(seed ^ splitmix64(a * 0x100000001B3ull + splitmix64(b)));
  return (double(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
inline double gauss(uint64_t seed, uint64_t a) {
  double u1 = u01(seed, a, 1), u2 = u01(seed, a, 2);
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
}
struct Box { double lo[3], hi[3]; };
struct Cyl { double cx, cy, r, h; };
struct Scene {
  uint64_t seed;
  int tx, ty;
  double tile;
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
};
}  // namespace

extern "C" {
void* synth_scene_create(uint64_t seed, int tiles_x, int tiles_y, double tile_size, int boxes_per_tile, int cyls_per_tile) {
  Scene* s = new Scene();
  s->seed = seed; s->tx = tiles_x; s->ty = tiles_y; s->tile = tile_size;
  for (int ix = 0; ix < tiles_x; ix++)
    for (int iy = 0; iy < tiles_y; iy++) {
      uint64_t tid = uint64_t(ix) * 4096 + uint64_t(iy);
      double ox = ix * tile_size, oy = iy * tile_size;
      for (int b = 0; b < boxes_per_tile; b++) {
        double w = 5 + 25 * u01(seed, tid, 100 + b * 8 + 0), d = 5 + 25 * u01(seed, tid, 100 + b * 8 + 1);
        double h = 3 + 17 * u01(seed, tid, 100 + b * 8 + 2);
        double cx = ox + tile_size * u01(seed, tid, 100 + b * 8 + 3), cy = oy + tile_size * u01(seed, tid, 100 + b * 8 + 4);
        Box bx{{cx - w / 2, cy - d / 2, 0.0}, {cx + w / 2, cy + d / 2, h}};
        s->boxes.push_back(bx);
      }
      for (int c = 0; c < cyls_per_tile; c++) {
        Cyl cy{ox + tile_size * u01(seed, tid, 5000 + c * 4 + 0), oy + tile_size * u01(seed, tid, 5000 + c * 4 + 1), 0.3, 8.0};
        s->cyls.push_back(cy);
      }
    }
  return s;
}
void synth_scene_destroy(void* h) { delete static_cast<Scene*>(h); }

// true if (x,y) at height z is strictly inside a box (used to keep sensor poses in free space)
int synth_point_free(void* h, double x, double y, double z, double margin) {
  Scene* s = static_cast<Scene*>(h);
  for (const Box& b : s->boxes)
    if (x > b.lo[0] - margin && x < b.hi[0] + margin && y > b.lo[1] - margin && y < b.hi[1] + margin && z < b.hi[2] + margin) return 0;
  for (const Cyl& c : s->cyls) {
    double dx = x - c.cx, dy = y - c.cy;
    if (dx * dx + dy * dy < (c.r + margin) * (c.r + margin) && z < c.h + margin) return 0;
  }
  return 1;
}

// Ray-cast one scan. T = sensor->world, column-major double[16]. Output points are in the SENSOR frame.
size_t synth_scan(void* h, const double* T, int beams, double el_min_deg, double el_max_deg, int az_steps, double max_range,
                  double noise_sigma, uint64_t noise_seed, float* out /* cap beams*az_steps*8 */) {
  Scene* s = static_cast<Scene*>(h);
  const double o[3] = {T[12], T[13], T[14]};
  // cull primitives by distance
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
  for (const Box& b : s->boxes) {
    double dx = std::max(std::max(b.lo[0] - o[0], 0.0), o[0] - b.hi[0]);
    double dy = std::max(std::max(b.lo[1] - o[1], 0.0), o[1] - b.hi[1]);
    if (dx * dx + dy * dy <= max_range * max_range) boxes.push_back(b);
  }
  for (const Cyl& c : s->cyls) {
    double dx = c.cx - o[0], dy = c.cy - o[1];
    if (std::sqrt(dx * dx + dy * dy) - c.r <= max_range) cyls.push_back(c);
  }
  const size_t nrays = size_t(beams) * az_steps;
  std::vector<float> ranges(nrays, -1.f);
  std::vector<float> dirs(nrays * 3);
#pragma omp parallel for schedule(dynamic, 512)
  for (long long ri = 0; ri < (long long)nrays; ri++) {
    int beam = int(ri / az_steps), az = int(ri % az_steps);
    double el = (beams > 1 ? el_min_deg + (el_max_deg - el_min_deg) * beam / double(beams - 1) : el_min_deg) * M_PI / 180.0;
    double a = 2.0 * M_PI * az / double(az_steps);
    double ds[3] = {std::cos(el) * std::cos(a), std::cos(el) * std::sin(a), std::sin(el)};
    double d[3];
    for (int r = 0; r < 3; r++) d[r] = T[r] * ds[0] + T[4 + r] * ds[1] + T[8 + r] * ds[2];
    double best = 1e30;
    if (d[2] < -1e-9) { double t = -o[2] / d[2]; if (t > 0) best = t; }
    for (const Box& b : boxes) {
      double t0 = 0, t1 = best;
      bool hit = true;
      for (int ax = 0; ax < 3 && hit; ax++) {
        if (std::fabs(d[ax]) < 1e-12) { if (o[ax] < b.lo[ax] || o[ax] > b.hi[ax]) hit = false; }
        else {
          double ta = (b.lo[ax] - o[ax]) / d[ax], tb = (b.hi[ax] - o[ax]) / d[ax];
          if (ta > tb) std::swap(ta, tb);
          t0 = std::max(t0, ta); t1 = std::min(t1, tb);
          if (t0 > t1) hit = false;
        }
      }
      if (hit && t0 > 1e-6 && t0 < best) best = t0;
    }
    for (const Cyl& c : cyls) {
      double ox = o[0] - c.cx, oy = o[1] - c.cy;
      double A = d[0] * d[0] + d[1] * d[1];
      if (A < 1e-12) continue;
      double B = ox * d[0] + oy * d[1], C = ox * ox + oy * oy - c.r * c.r;
      double disc = B * B - A * C;
      if (disc < 0) continue;
      double t = (-B - std::sqrt(disc)) / A;
      if (t > 1e-6 && t < best) { double z = o[2] + t * d[2]; if (z >= 0 && z <= c.h) best = t; }
    }
    if (best < 1e29) {
      double rng = best + noise_sigma * gauss(noise_seed, uint64_t(ri));
      if (rng > 0.5 && rng <= max_range) {
        ranges[ri] = float(rng);
        dirs[ri * 3] = float(ds[0] * rng); dirs[ri * 3 + 1] = float(ds[1] * rng); dirs[ri * 3 + 2] = float(ds[2] * rng);
      }
    }
  }
  size_t n = 0;
  for (size_t ri = 0; ri < nrays; ri++) {
    if (ranges[ri] < 0) continue;
    float* p = out + n * 8;
    p[0] = dirs[ri * 3]; p[1] = dirs[ri * 3 + 1]; p[2] = dirs[ri * 3 + 2]; p[3] = 1.f;
    p[4] = 0.f; p[5] = p[6] = p[7] = 0.f;
    n++;
  }
  return n;
}

// Direct surface sampler (jittered lattice of the given spacing + Gaussian noise along the surface normal) over
// the world-frame rectangle [x0,x1]x[y0,y1]. Pass out = nullptr to count. Deterministic for a given seed.
size_t synth_sample_map(void* h, double x0, double y0, double x1, double y1, double spacing, double noise_sigma, uint64_t seed,
                        float* out, size_t cap) {
  Scene* s = static_cast<Scene*>(h);
  size_t n = 0;
  auto emit = [&](double x, double y, double z) {
    if (x < x0 || x >= x1 || y < y0 || y >= y1) return;
    if (out && n < cap) { float* p = out + n * 8; p[0] = float(x); p[1] = float(y); p[2] = float(z); p[3] = 1.f; p[4] = p[5] = p[6] = p[7] = 0.f; }
    n++;
  };
  uint64_t ctr = 0;
  // ground (skip footprints of boxes)
  long long gx = (long long)std::ceil((x1 - x0) / spacing), gy = (long long)std::ceil((y1 - y0) / spacing);
  for (long long iy = 0; iy < gy; iy++)
    for (long long ix = 0; ix < gx; ix++) {
      uint64_t id = uint64_t(iy) * uint64_t(gx) + uint64_t(ix);
      double x = x0 + (ix + u01(seed, id, 11)) * spacing, y = y0 + (iy + u01(seed, id, 12)) * spacing;
      bool covered = false;
      // coarse: check boxes (few thousand at most; fine for the sizes used)
      for (const Box& b : s->boxes) if (x > b.lo[0] && x < b.hi[0] && y > b.lo[1] && y < b.hi[1]) { covered = true; break; }
      if (covered) continue;
      emit(x, y, noise_sigma * gauss(seed, id * 3 + 1));
    }
  ctr = uint64_t(gx) * uint64_t(gy) * 4;
  for (size_t bi = 0; bi < s->boxes.size(); bi++) {
    const Box& b = s->boxes[bi];
    if (b.hi[0] < x0 || b.lo[0] > x1 || b.hi[1] < y0 || b.lo[1] > y1) continue;
    // 4 walls + roof
    for (int w = 0; w < 4; w++) {
      double len = (w < 2) ? (b.hi[0] - b.lo[0]) : (b.hi[1] - b.lo[1]);
      long long nu = (long long)std::ceil(len / spacing), nv = (long long)std::ceil(b.hi[2] / spacing);
      for (long long iu = 0; iu < nu; iu++)
        for (long long iv = 0; iv < nv; iv++) {
          uint64_t id = ctr + (uint64_t(bi) * 8 + w) * 1000003ull + uint64_t(iu) * 4099 + uint64_t(iv);
          double u = (iu + u01(seed, id, 21)) * spacing, v = (iv + u01(seed, id, 22)) * spacing;
          if (u > len || v > b.hi[2]) continue;
          double nz = noise_sigma * gauss(seed, id * 3 + 2);
          if (w == 0) emit(b.lo[0] + u, b.lo[1] + nz, v);
          else if (w == 1) emit(b.lo[0] + u, b.hi[1] + nz, v);
          else if (w == 2) emit(b.lo[0] + nz, b.lo[1] + u, v);
          else emit(b.hi[0] + nz, b.lo[1] + u, v);
        }
    }
    long long nu = (long long)std::ceil((b.hi[0] - b.lo[0]) / spacing), nv = (long long)std::ceil((b.hi[1] - b.lo[1]) / spacing);
    for (long long iu = 0; iu < nu; iu++)
      for (long long iv = 0; iv < nv; iv++) {
        uint64_t id = ctr + (uint64_t(bi) * 8 + 5) * 1000003ull + uint64_t(iu) * 4099 + uint64_t(iv);
        double u = (iu + u01(seed, id, 23)) * spacing, v = (iv + u01(seed, id, 24)) * spacing;
        if (u > b.hi[0] - b.lo[0] || v > b.hi[1] - b.lo[1]) continue;
        emit(b.lo[0] + u, b.lo[1] + v, b.hi[2] + noise_sigma * gauss(seed, id * 3 + 2));
      }
  }
  for (size_t ci = 0; ci < s->cyls.size(); ci++) {
    const Cyl& c = s->cyls[ci];
    if (c.cx < x0 - 1 || c.cx > x1 + 1 || c.cy < y0 - 1 || c.cy > y1 + 1) continue;
    long long na = std::max(3LL, (long long)std::ceil(2 * M_PI * c.r / spacing)), nv = (long long)std::ceil(c.h / spacing);
    for (long long ia = 0; ia < na; ia++)
      for (long long iv = 0; iv < nv; iv++) {
        uint64_t id = ctr + (1ull << 40) + uint64_t(ci) * 100003ull + uint64_t(ia) * 1031 + uint64_t(iv);
        double a = 2 * M_PI * (ia + u01(seed, id, 31)) / double(na), v = (iv + u01(seed, id, 32)) * spacing;
        if (v > c.h) continue;
        double r = c.r + noise_sigma * gauss(seed, id * 3 + 2);
        emit(c.cx + r * std::cos(a), c.cy + r * std::sin(a), v);
      }
  }
  return n;
}
}
