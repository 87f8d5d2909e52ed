// Timeline reconstruction: per-chunk speaker activity + per-chunk cluster assignments -> speaker segments.
// Reference: Sources/FluidAudio/Diarizer/Offline/Utils/OfflineReconstruction.swift:24-253 (buildSegments),
//            :359-505 (excludeOverlaps, appendSegment, mergeSegments, blendedQuality, sanitize, chunkStartTime).
// The step after buildChunkAssignments in OfflineDiarizerManager.cluster(_:) (:418-424).  Its inputs come from the
// segmentation model (not re-implemented): weights [chunks x frames x local speakers].  O(chunks x frames x speakers)
// additions followed by a sequential sweep over the global frames: host work, like the reference.
#pragma once
#include <cstdint>
#include <vector>

namespace fa {
namespace reconstruct {

struct Config {
    double frame_duration = 0.0;        // SegmentationOutput.frameDuration
    double window_duration = 10.0;      // config.windowDuration: chunk c starts at c * window_duration without offsets
    double min_gap_duration = 0.1;      // PostProcessing.minGapDurationSeconds
    double seg_min_duration_off = 0.0;  // Segmentation.minDurationOff
    double seg_min_duration_on = 0.0;   // Segmentation.minDurationOn
    double min_segment_duration = 1.0;  // Embedding.minSegmentDurationSeconds
    bool exclusive_segments = true;     // PostProcessing.exclusiveSegments
};

struct Segment {
    int32_t cluster;   // speakerId = "S\(cluster + 1)", embedding = centroids[cluster]
    float start, end, quality;
};

void build_segments(const float *weights, int num_chunks, int num_frames, int num_speakers, const double *chunk_offsets,
                    int offsets_count, const int32_t *hard_clusters, int hard_rows, int centroid_count, const Config &cfg,
                    std::vector<Segment> &out);

// buildSpeakerDatabase (:296-357): float32 mean of the segment embeddings (= Float(centroid)) per speaker.
void build_speaker_database(const int32_t *seg_cluster, int seg_count, const double *centroids, int K, int dim,
                            float *database, int32_t *counts);

} // namespace reconstruct
} // namespace fa
