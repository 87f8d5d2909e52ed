// Host-side solver for centroid-linkage agglomerative clustering on one B200 (see ahc_kernels.cu).
#pragma once

#include "ahc_core.cuh"
#include <cuda_runtime.h>
#include <vector>

namespace fa {
namespace ahc {

// One command from the master thread to the worker CTAs of the persistent merge kernel.
struct Command {
    int type;        // 1 = MERGE (build node `fresh` from a,b then scan), 2 = RESCAN (scan for node `target`), 3 = EXIT
    int a, b;        // node ids being merged
    int fresh;       // id of the new node (MERGE) / target node id (RESCAN)
    int slot_a;      // slot holding a: becomes the slot of `fresh`
    int slot_b;      // slot holding b: becomes empty
    int limit;       // only nodes with id < limit are candidates
    int pad;
    double wa, wb;   // member counts of a and b as doubles
};

// Everything the persistent kernel needs, resident in HBM.
struct Problem {
    int N, D, Ns;            // points, dimension, slot stride (N rounded up to 32)
    double *rows;            // [(2N-1) x D] node store, row-major: rows 0..N-1 = input, N.. = merged centroids
    double *cols;            // [D x Ns]     scan copy, k-major: cols[k*Ns + slot]
    // master state
    double *key;             // [2N-2] nearest-neighbour distance per node id (heap keys)
    int *nn;                 // [2N-2] nearest neighbour per node id
    int *heap_at;            // [N-1]
    int *heap_where;         // [2N-2]
    int *live_next;          // [2N]
    int *live_prev;          // [2N]
    int *weight;             // [2N-1] member count per node
    int *slot_of;            // [2N-1] slot per node
    int *merge_a, *merge_b;  // [N-1] merge log
    double *merge_d;         // [N-1] squared distance of each merge
    // synchronisation
    Command *cmd;            // 1
    unsigned *seq;           // command sequence number (release/acquire)
    unsigned *arrive;        // worker arrival counter (monotonic)
    Cand *partial;           // [workers] per-CTA minima
    int *error;              // 0 ok, 1 NaN distance
    int heap_size;           // after host heapify
    int steps_done;          // (debug) merges completed
};

struct Solver {
    int num_sms = 0;
    int max_workers = 0;      // worker CTAs available for one problem (grid = workers + 1)
    long long launches = 0;
    cudaStream_t stream = nullptr;
    // device storage (grown on demand, reused across calls)
    void *d_pool = nullptr;
    size_t pool_bytes = 0;
    Problem *d_problem = nullptr;
    double *d_input = nullptr;   // staging of the caller's [N x D] rows when they come from the host
    size_t input_cap = 0;
    // pinned host mirrors
    void *h_pool = nullptr;
    size_t h_pool_bytes = 0;
    float last_ms[4] = {0, 0, 0, 0};   // init-nn, host heapify + copies, merge loop, total

    ~Solver();
    void release();
    int init(cudaStream_t s, int worker_limit);
    // rows: device pointer to N x D row-major doubles (already normalised by the caller, as the reference requires).
    // Z: host buffer of (N-1) x 4 doubles.  Status codes follow FastClusterWrapper.h.
    int linkage_device(const double *d_rows, int N, int D, double *Z_host);
    int linkage_host(const double *rows_host, size_t N, size_t D, double *Z_host, size_t z_len);
    int ensure_pool(int N, int D);
};

// Standalone kernels used by the clustering pipeline
int launch_normalize_rows(const double *d_in, double *d_out, int rows, int dim, cudaStream_t s);
int launch_widen_rows(const float *d_in, double *d_out, long long count, cudaStream_t s);

// Swift-side dendrogram cut + first-appearance relabel (AHCClustering.swift:112-121,124-210), host, O(N).
void dendrogram_cut(const double *Z, long long count, double threshold, int32_t *labels);

} // namespace ahc
} // namespace fa
