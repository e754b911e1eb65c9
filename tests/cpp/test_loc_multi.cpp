// loc.cpp-style host (reference: test/loc.cpp:47-63 — a static map, then one scan2Map per incoming scan) using several
// GPUs from ONE C++ process through pcr_multi_* : the index is built once, copied to the peers inside the library, the
// scans are sharded in contiguous blocks. Plain C ABI, no Python, no torch.
//   test_loc_multi <method: loam|ndt> <n_devices> [same]   ("same" = every context on device 0: exercises the path on a 1-GPU box)
// Exit codes: 0 ok, 3 no CUDA device (expected on the CPU-only box), 1 wrong result.
#include <pcr_cuda.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct P32 { float x, y, z, one, intensity, pad[3]; };

static void make_room(std::vector<P32>& pc) {  // 24 x 18 x 5 m room sampled every 0.12 m with a deterministic ripple
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double(s >> 8) / double(1u << 24) - 0.5) * 0.01; };
  auto add = [&](double x, double y, double z) { P32 p{}; p.x = float(x); p.y = float(y); p.z = float(z); p.one = 1.f; pc.push_back(p); };
  for (double x = -12; x <= 12; x += 0.12)
    for (double y = -9; y <= 9; y += 0.12) add(x + rnd(), y + rnd(), rnd());
  for (double z = 0; z <= 5; z += 0.12) {
    for (double x = -12; x <= 12; x += 0.12) { add(x + rnd(), -9 + rnd(), z); add(x + rnd(), 9 + rnd(), z); }
    for (double y = -9; y <= 9; y += 0.12) { add(-12 + rnd(), y + rnd(), z); add(12 + rnd(), y + rnd(), z); }
  }
}

int main(int argc, char** argv) {
  const char* method = argc > 1 ? argv[1] : "loam";
  const int nd = argc > 2 ? std::atoi(argv[2]) : 2;
  const bool same = argc > 3 && std::strcmp(argv[3], "same") == 0;
  pcr_params prm;
  pcr_default_params(std::strcmp(method, "ndt") == 0 ? PCR_NDT : PCR_LOAM, &prm);
  std::vector<int32_t> devs;
  for (int k = 0; k < nd; k++) devs.push_back(same ? 0 : k);
  pcr_multi* m = nullptr;
  int rc = pcr_multi_create(&prm, devs.data(), devs.size(), &m);
  if (rc == PCR_ERR_NO_DEVICE) { std::printf("no device: %s\n", pcr_last_error(nullptr)); return 3; }
  if (rc) { std::printf("create failed: %s\n", pcr_last_error(nullptr)); return 1; }
  std::vector<P32> map;
  make_room(map);
  rc = pcr_multi_set_target(m, map.data(), map.size(), sizeof(P32));
  if (rc) { std::printf("set_target failed: %s\n", pcr_multi_last_error(m)); return 1; }
  size_t blob = 0; double ms = 0;
  pcr_multi_get_broadcast(m, &blob, &ms);
  // 11 scans: every 3rd (+k) map point seen from a slightly different pose each; guess = identity
  const int n_scans = 11;
  std::vector<P32> scans;
  std::vector<size_t> offs(1, 0);
  std::vector<double> truth;
  for (int k = 0; k < n_scans; k++) {
    const double yaw = (1.0 + 0.2 * k) * M_PI / 180.0, tx = 0.2 - 0.03 * k, ty = -0.15 + 0.02 * k, tz = 0.04;
    for (size_t i = size_t(k % 3); i < map.size(); i += 3) {
      const double dx = map[i].x - tx, dy = map[i].y - ty, dz = map[i].z - tz;
      P32 p{}; p.one = 1.f;
      p.x = float(std::cos(yaw) * dx + std::sin(yaw) * dy); p.y = float(-std::sin(yaw) * dx + std::cos(yaw) * dy); p.z = float(dz);
      scans.push_back(p);
    }
    offs.push_back(scans.size());
    truth.push_back(tx); truth.push_back(ty); truth.push_back(tz); truth.push_back(yaw);
  }
  std::vector<double> T(16 * n_scans, 0.0);
  for (int k = 0; k < n_scans; k++) for (int d = 0; d < 4; d++) T[16 * k + 5 * d] = 1.0;
  std::vector<int32_t> conv(n_scans, 0);
  rc = pcr_multi_batch_align(m, scans.data(), offs.data(), n_scans, sizeof(P32), T.data(), conv.data());
  if (rc) { std::printf("batch_align failed: %s\n", pcr_multi_last_error(m)); return 1; }
  // the same job through ONE context must give the same poses
  pcr_ctx* c = nullptr;
  prm.device = 0;
  if (pcr_create(&prm, &c) || pcr_set_target(c, map.data(), map.size(), sizeof(P32))) { std::printf("single ctx failed\n"); return 1; }
  std::vector<double> T1(16 * n_scans, 0.0);
  for (int k = 0; k < n_scans; k++) for (int d = 0; d < 4; d++) T1[16 * k + 5 * d] = 1.0;
  std::vector<int32_t> conv1(n_scans, 0);
  if (pcr_batch_align(c, scans.data(), offs.data(), n_scans, sizeof(P32), T1.data(), conv1.data())) { std::printf("single batch failed: %s\n", pcr_last_error(c)); return 1; }
  int bad = 0, nconv = 0;
  const bool ndt = prm.method == PCR_NDT;
  for (int k = 0; k < n_scans; k++) {
    const double* P = &T[16 * k];
    const double et = std::sqrt(std::pow(P[12] - truth[4 * k], 2) + std::pow(P[13] - truth[4 * k + 1], 2) + std::pow(P[14] - truth[4 * k + 2], 2));
    const double er = std::fabs(std::atan2(P[1], P[0]) - truth[4 * k + 3]);
    double dmax = 0;
    for (int q = 0; q < 16; q++) dmax = std::fmax(dmax, std::fabs(P[q] - T1[16 * k + q]));
    nconv += conv[k];
    if (!(et < (ndt ? 0.15 : 0.03)) || !(er < (ndt ? 0.02 : 0.006)) || conv[k] != conv1[k] || !(dmax < 1e-6)) {
      std::printf("scan %d: t_err %.4f r_err %.5f conv %d/%d max |multi - single| %.3g\n", k, et, er, conv[k], conv1[k], dmax);
      bad++;
    }
  }
  std::printf("%s on %d context(s)%s: %d scans, %d converged, index blob %zu bytes copied to each peer in %.3f ms, %d bad\n", method, nd,
              same ? " (same device)" : "", n_scans, nconv, blob, ms, bad);
  pcr_destroy(c);
  pcr_multi_destroy(m);
  return bad ? 1 : 0;
}
