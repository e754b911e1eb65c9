// Exercises the C++ adaptor (PCR::LoamRegister / NdtRegister / VgicpRegister over the C ABI) the way
// test/align.cpp:111-146 of the reference drives a register: makeRegister(name) -> scan2Map(src, dst, pose).
// Exit codes: 0 ok, 3 no CUDA device (expected on the CPU-only box), 1 wrong result.
#include <PCR/LoamRegister.hpp>
#include <PCR/NdtRegister.hpp>
#include <PCR/VgicpRegister.hpp>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static pc_t::Ptr make_room(double noise_seed) {
  // a 20 x 16 x 5 m room sampled every 0.12 m (floor + 4 walls) with a small deterministic ripple
  auto pc = std::make_shared<pc_t>();
  unsigned s = unsigned(noise_seed * 7919) + 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double(s >> 8) / double(1u << 24) - 0.5) * 0.01; };
  for (double x = -10; x <= 10; x += 0.12)
    for (double y = -8; y <= 8; y += 0.12) { pt_t p; p.x = float(x + rnd()); p.y = float(y + rnd()); p.z = float(rnd()); pc->push_back(p); }
  for (double z = 0; z <= 5; z += 0.12) {
    for (double x = -10; x <= 10; x += 0.12) {
      pt_t p; p.x = float(x + rnd()); p.y = float(-8 + rnd()); p.z = float(z); pc->push_back(p);
      p.y = float(8 + rnd()); pc->push_back(p);
    }
    for (double y = -8; y <= 8; y += 0.12) {
      pt_t p; p.x = float(-10 + rnd()); p.y = float(y + rnd()); p.z = float(z); pc->push_back(p);
      p.x = float(10 + rnd()); pc->push_back(p);
    }
  }
  return pc;
}

int main(int argc, char** argv) {
  const char* names[3] = {"loam", "ndt", "vgicp"};
  try {
    PCR::makeRegister("icp");
    std::printf("unknown register name did not throw\n");
    return 1;
  } catch (const std::runtime_error&) {
  }
  auto dst = make_room(1.0);
  // source = every 3rd map point moved by the inverse of a known pose (yaw 2 deg, t = (0.25, -0.15, 0.05))
  const double yaw = 2.0 * M_PI / 180.0, tx = 0.25, ty = -0.15, tz = 0.05;
  auto src = std::make_shared<pc_t>();
  for (size_t i = 0; i < dst->size(); i += 3) {
    const pt_t& m = dst->points[i];
    const double dx = m.x - tx, dy = m.y - ty, dz = m.z - tz;
    pt_t p;
    p.x = float(std::cos(yaw) * dx + std::sin(yaw) * dy);
    p.y = float(-std::sin(yaw) * dx + std::cos(yaw) * dy);
    p.z = float(dz);
    src->push_back(p);
  }
  int bad = 0;
  for (int k = 0; k < 3; k++) {
    if (argc > 1 && std::strcmp(argv[1], names[k]) != 0) continue;
    PCR::PointCloudRegister::Ptr reg;
    try {
      reg = PCR::makeRegister(names[k]);
    } catch (const std::runtime_error& e) {
      std::printf("%s: construction failed: %s\n", names[k], e.what());
      return 3;
    }
    pose_t pose;  // identity guess
    const bool conv = reg->scan2Map(src, dst, pose);
    const double ex = pose.matrix()(0, 3) - tx, ey = pose.matrix()(1, 3) - ty, ez = pose.matrix()(2, 3) - tz;
    const double eyaw = std::atan2(pose.matrix()(1, 0), pose.matrix()(0, 0)) - yaw;
    const double et = std::sqrt(ex * ex + ey * ey + ez * ez);
    std::printf("%-5s converged=%d  t_err=%.4f m  yaw_err=%.5f rad  fitness=%.5f\n", names[k], int(conv), et, eyaw, reg->getFitnessScore());
    const double tol_t = (k == 1) ? 0.15 : 0.03, tol_r = (k == 1) ? 0.02 : 0.004;  // NDT stops at its 0.1 step epsilon
    if (!(et < tol_t) || !(std::fabs(eyaw) < tol_r)) bad++;
    // localisation mode: target cached by (pointer, size, generation) and its points page-locked for the upload; the second
    // call reuses the index, a bumped generation rebuilds it — same pose every time
    auto* cr = dynamic_cast<PCR::CudaRegister*>(reg.get());
    if (cr) {
      cr->enableTargetCache(true, true);
      pose_t p1, p2, p3;
      const bool c1 = reg->scan2Map(src, dst, p1), c2 = reg->scan2Map(src, dst, p2);
      cr->bumpTargetGeneration();
      const bool c3 = reg->scan2Map(src, dst, p3);
      bool same = c1 == conv && c2 == conv && c3 == conv;
      for (int q = 0; q < 16; q++) same = same && p1.matrix().data()[q] == pose.matrix().data()[q] && p2.matrix().data()[q] == p1.matrix().data()[q] &&
                                         p3.matrix().data()[q] == p1.matrix().data()[q];
      std::printf("%-5s target cache + pinned target: %s\n", names[k], same ? "same pose" : "MISMATCH");
      if (!same) bad++;
      cr->enableTargetCache(false);
    }
  }
  return bad ? 1 : 0;
}
