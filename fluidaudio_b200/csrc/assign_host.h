// Per-chunk constrained assignment (host side of the clustering path).
// Reference: Sources/FluidAudio/Diarizer/HungarianAssignment.swift:8-97,
//            Sources/FluidAudio/Diarizer/Offline/Clustering/ConstrainedClusterAssignment.swift:20-42,
//            Sources/FluidAudio/Diarizer/Offline/Core/OfflineDiarizerManager.swift:885-911.
// A chunk holds at most a handful of local speakers (<= 3 in the reference's segmentation model) and K clusters, so
// each problem is a <= max(3, K)^2 integer matching: O(chunks * K^3) exact integer work, done on the host between
// the GPU's N x K cosine-score kernel and the label copy-out.  There is nothing data-parallel to accelerate here.
#pragma once
#include <cstdint>

namespace fa {
namespace assign {

// Minimum-cost perfect matching on an n x n non-negative integer matrix (row-major); out[row] = column.
void min_cost_matching(const int64_t *cost, int n, int32_t *out);
// Maximum-total-score matching on a rows x cols score matrix; out[row] = column or -1 (more rows than columns).
void max_score_matching(const double *scores, int rows, int cols, int32_t *out);
// scores: N x K, chunk[N]; out[N] = cluster, or -2 where a chunk has more local speakers than clusters.
void constrained_assign(const double *scores, long long N, int K, const int32_t *chunk, int32_t *out);
// [num_chunks x num_speakers] matrix of cluster ids, -2 where nothing valid was assigned.
void build_chunk_assignments(const int32_t *chunk, const int32_t *speaker, const int32_t *assignments, long long N,
                             int num_chunks, int num_speakers, int cluster_count, int32_t *matrix);

} // namespace assign
} // namespace fa
