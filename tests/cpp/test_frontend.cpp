// Drives the C++ headless frontend (frontend::LidarOdometry over the C ABI) on a frame file written by
// tests/test_cpp_frontend.py and writes the per-frame global poses back; the Python test compares them with the
// ctypes mirror (simpleslam_b200/frontend.py) and with the oracle.
// frame file: int64 n_frames, then per frame: double stamp, double local_odom[16] (column-major), int64 n_pts, n_pts * 8 floats
// output: per frame 16 doubles (column-major) + int64 converged; then int64 n_keyframes, int64 submap_size
// exit codes: 0 ok, 3 no CUDA device, 2 usage / io
#include <frontend/HeadlessOdometry.hpp>
#include <cstdio>
#include <cstdint>

int main(int argc, char** argv) {
  if (argc < 4) { std::printf("usage: test_frontend <pcr> <frames.bin> <poses.bin>\n"); return 2; }
  std::unique_ptr<frontend::LidarOdometry> lo;
  try {
    lo.reset(new frontend::LidarOdometry(argv[1]));
  } catch (const std::runtime_error& e) {
    std::printf("construction failed: %s\n", e.what());
    return 3;
  }
  FILE* in = std::fopen(argv[2], "rb");
  FILE* out = std::fopen(argv[3], "wb");
  if (!in || !out) return 2;
  int64_t n = 0;
  if (std::fread(&n, 8, 1, in) != 1) return 2;
  for (int64_t k = 0; k < n; k++) {
    double stamp;
    pose_t odom;
    int64_t np = 0;
    if (std::fread(&stamp, 8, 1, in) != 1 || std::fread(odom.matrix().data(), 8, 16, in) != 16 || std::fread(&np, 8, 1, in) != 1) return 2;
    auto pc = std::make_shared<pc_t>();
    pc->points.resize(size_t(np));
    if (np && std::fread(pc->points.data(), 32, size_t(np), in) != size_t(np)) return 2;
    const pose_t P = lo->generateOdom(pc, stamp, &odom);
    const int64_t conv = lo->lastConverged() ? 1 : 0;
    std::fwrite(P.matrix().data(), 8, 16, out);
    std::fwrite(&conv, 8, 1, out);
  }
  const int64_t nk = int64_t(lo->map().keyframes().size()), sm = int64_t(lo->map().submapSize());
  std::fwrite(&nk, 8, 1, out);
  std::fwrite(&sm, 8, 1, out);
  std::fclose(in);
  std::fclose(out);
  std::printf("frames %lld keyframes %lld submap %lld\n", (long long)n, (long long)nk, (long long)sm);
  return 0;
}
