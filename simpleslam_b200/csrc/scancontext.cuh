// ScanContext place-recognition descriptor on the GPU: backend/src/ScanContext.cpp:152-233
// (makeScanContext, ring / sector keys) and :73-150 (fastAlignUsingVkey, computeSimularity, distanceBtnScanContext).
#pragma once
#include "common.cuh"

namespace pcr {

constexpr int kScRings = 20;     // ScanContext::PC_NUM_RING
constexpr int kScSectors = 60;   // ScanContext::PC_NUM_SECTOR
constexpr float kScMaxRadius = 80.0f;  // ScanContext::PC_MAX_RADIUS
constexpr int kScSize = kScRings * kScSectors;

// one descriptor per cloud: clouds are concatenated float4 points, cloud i = [offs[i], offs[i+1]).
// desc: n_clouds x 1200 doubles (row-major [ring][sector]); ring_key: n_clouds x 20; sector_key: n_clouds x 60 (device).
void scancontext_make(const float4* pts, const uint32_t* d_offs, int n_clouds, float lidar_height, double* desc, double* ring_key,
                      double* sector_key, cudaStream_t s);
// distanceBtnScanContext for a batch of descriptor pairs (device arrays): dist[p], shift[p]
// key_align: 0 = alignment shift fixed at 0 (what the reference computes, see the kernel), 1 = all 60 shifts of the sector key
void scancontext_distance(const double* desc, const int2* pairs, int n_pairs, int search_radius, int key_align, double* dist, int32_t* shift,
                          cudaStream_t s);

}  // namespace pcr
