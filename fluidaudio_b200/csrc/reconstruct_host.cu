// See reconstruct_host.h.  Written around flat per-frame tables (one pass to aggregate, one sweep to cut segments);
// every comparison and rounding follows the Swift source cited at each step.
#include "reconstruct_host.h"

#include <algorithm>
#include <cmath>

namespace fa {
namespace reconstruct {

namespace {

struct Open {          // a segment being extended (OfflineReconstruction.Accumulator :8-13)
    double start, end, score;
    int frames;
    bool on;
};

float clamp01f(double v) { return (float)std::min(std::max(v, 0.0), 1.0); }

float blend(const Segment &l, const Segment &r) {   // blendedQuality :462-476
    const double ld = (double)(l.end - l.start), rd = (double)(r.end - r.start), total = ld + rd;
    if (!(total > 0.0)) return std::min(std::max((l.quality + r.quality) / 2, 0.0f), 1.0f);
    return clamp01f(((double)l.quality * ld + (double)r.quality * rd) / total);
}

void sort_by_start(std::vector<Segment> &v) {        // Swift's sorted(by:) is stable
    std::stable_sort(v.begin(), v.end(), [](const Segment &a, const Segment &b) { return a.start < b.start; });
}

} // namespace

void build_segments(const float *weights, int num_chunks, int num_frames, int num_speakers, const double *chunk_offsets,
                    int offsets_count, const int32_t *hard_clusters, int hard_rows, int centroid_count, const Config &cfg,
                    std::vector<Segment> &out) {
    out.clear();
    const double dur = cfg.frame_duration;
    if (num_chunks <= 0 || num_frames <= 0 || !(dur > 0.0)) return;                       // :29-32
    const int K = std::max(centroid_count, 1);                                            // :34
    const double gap_threshold = std::max(cfg.min_gap_duration, cfg.seg_min_duration_off);   // :35
    auto offset_of = [&](int c) { return c < offsets_count ? chunk_offsets[c] : (double)c * cfg.window_duration; };   // :495-505

    double latest = 0.0;                                                                  // :37-44
    for (int c = 0; c < num_chunks; ++c) latest = std::max(latest, offset_of(c) + (double)num_frames * dur);
    const int G = std::max(1, (int)std::ceil(latest / dur));                              // :46

    // ---- aggregation over chunks (:58-105): per global frame and cluster the sum / count of the strongest local
    //      speaker mapped to that cluster, plus the expected number of simultaneous speakers
    std::vector<double> vote((size_t)G * K, 0.0), votes((size_t)G * K, 0.0), expect(G, 0.0), seen(G, 0.0), local(K);
    for (int c = 0; c < num_chunks; ++c) {
        const double off = offset_of(c);
        const int32_t *map = c < hard_rows ? hard_clusters + (size_t)c * num_speakers : nullptr;
        for (int f = 0; f < num_frames; ++f) {
            int g = (int)std::round((off + (double)f * dur) / dur);                       // .rounded(): ties away from zero (:70)
            g = g < 0 ? 0 : (g >= G ? G - 1 : g);
            const float *w = weights + ((size_t)c * num_frames + f) * num_speakers;
            std::fill(local.begin(), local.end(), 0.0);
            double sum_w = 0.0;
            for (int s = 0; s < num_speakers; ++s) {
                const double v = (double)w[s];
                sum_w += v;                                                               // expectedCount (:90-92), every speaker
                const int k = map ? map[s] : -2;
                if (k >= 0 && k < K && v > local[k]) local[k] = v;                        // :81-88
            }
            expect[g] += sum_w;
            seen[g] += 1.0;
            double *vs = &vote[(size_t)g * K], *vc = &votes[(size_t)g * K];
            for (int k = 0; k < K; ++k)
                if (local[k] > 0.0) {                                                     // :96-102
                    vs[k] += local[k];
                    vc[k] += 1.0;
                }
        }
    }

    // ---- one sweep over the global frames: speaker count (:145-157), top clusters by vote sum (:168-175), segment
    //      accumulation (:188-235)
    const int most = std::min(K, num_speakers);
    std::vector<Open> open(K, Open{0, 0, 0, 0, false});
    std::vector<int> rank(K);
    std::vector<char> chosen(K);
    std::vector<Segment> raw;
    auto close = [&](int k, double end_time) {                                            // appendSegment :399-425
        const Open &o = open[k];
        if (end_time > o.start) {
            const double mean = o.frames > 0 ? o.score / (double)o.frames : o.score;
            raw.push_back(Segment{k, (float)o.start, (float)end_time, clamp01f(mean)});
        }
    };
    for (int g = 0; g < G; ++g) {
        std::fill(chosen.begin(), chosen.end(), 0);
        if (seen[g] > 0.0) {
            int need = (int)std::nearbyint(expect[g] / seen[g]);                          // .rounded(.toNearestOrEven)
            need = need < 0 ? 0 : (need > most ? most : need);
            if (need > 0) {
                const double *vs = &vote[(size_t)g * K];
                for (int k = 0; k < K; ++k) rank[k] = k;
                std::stable_sort(rank.begin(), rank.end(), [&](int a, int b) { return vs[a] > vs[b]; });
                for (int i = 0; i < need; ++i) chosen[rank[i]] = 1;
            }
        }
        const double t0 = (double)g * dur, t1 = t0 + dur;
        for (int k = 0; k < K; ++k)                      // clusters in ascending order (the reference iterates a Dictionary)
            if (open[k].on && !chosen[k]) {
                close(k, t0);
                open[k].on = false;
            }
        for (int k = 0; k < K; ++k) {
            if (!chosen[k]) continue;
            const double vc = votes[(size_t)g * K + k];
            const double score = vc == 0.0 ? 0.0 : vote[(size_t)g * K + k] / vc;          // activationAverages :107-142
            Open &o = open[k];
            if (o.on) {
                o.end = t1;
                o.score += score;
                o.frames += 1;
            } else {
                o = Open{t0, t1, score, 1, true};
            }
        }
    }
    for (int k = 0; k < K; ++k)
        if (open[k].on) close(k, open[k].end);                                            // :237-245

    // ---- mergeSegments (:427-460)
    std::vector<Segment> merged;
    if (!raw.empty()) {
        sort_by_start(raw);
        Segment cur = raw.front();
        for (size_t i = 1; i < raw.size(); ++i) {
            const Segment &nx = raw[i];
            if (nx.cluster == cur.cluster && (double)nx.start - (double)cur.end <= gap_threshold) {
                const float q = blend(cur, nx);
                cur.end = std::max(cur.end, nx.end);
                cur.quality = q;
            } else {
                merged.push_back(cur);
                cur = nx;
            }
        }
        merged.push_back(cur);
    }
    // ---- sanitize (:478-493) and excludeOverlaps (:359-397)
    sort_by_start(merged);
    const float keep_from = std::max((float)cfg.min_segment_duration, (float)cfg.seg_min_duration_on);
    for (const Segment &sg : merged) {
        if (!(sg.end - sg.start >= keep_from)) continue;
        if (!cfg.exclusive_segments) {
            out.push_back(sg);
            continue;
        }
        float start = sg.start;
        if (!out.empty() && start < out.back().end) start = out.back().end;
        if (start >= sg.end) continue;
        const float len = sg.end - start;
        if (len < (float)cfg.min_segment_duration) continue;
        const float whole = sg.end - sg.start;
        const float scale = whole > 0 ? len / whole : 1.0f;
        out.push_back(Segment{sg.cluster, start, sg.end, std::max(0.0f, std::min(1.0f, sg.quality * scale))});
    }
}

void build_speaker_database(const int32_t *seg_cluster, int seg_count, const double *centroids, int K, int dim,
                            float *database, int32_t *counts) {
    std::fill(counts, counts + K, 0);
    std::fill(database, database + (size_t)K * dim, 0.0f);
    for (int s = 0; s < seg_count; ++s) {
        const int k = seg_cluster[s];
        if (k < 0 || k >= K) continue;                       // a cluster without centroid contributes a zero embedding
        float *row = database + (size_t)k * dim;
        const double *c = centroids + (size_t)k * dim;
        if (counts[k] == 0) {
            for (int q = 0; q < dim; ++q) row[q] = (float)c[q];           // sums[speaker] = segment.embedding (:330)
        } else {
            for (int q = 0; q < dim; ++q) row[q] = row[q] + (float)c[q];  // cblas_saxpy, alpha = 1 (:313-325)
        }
        ++counts[k];
    }
    for (int k = 0; k < K; ++k) {
        if (counts[k] <= 0) continue;
        const float scale = 1.0f / (float)counts[k];                     // vDSP_vsmul (:341-352)
        float *row = database + (size_t)k * dim;
        for (int q = 0; q < dim; ++q) row[q] *= scale;
    }
}

} // namespace reconstruct
} // namespace fa
