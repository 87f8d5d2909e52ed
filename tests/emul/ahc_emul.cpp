// Host emulator of the merge kernel's CONTROL PLANE (CPU test-suite only).
// Drives the exact NnHeapT<uint16_t> / LiveSet code the device master warp runs (fluidaudio_b200/csrc/ahc_core.cuh),
// slot-indexed, in the same order as ahc_master() in ahc_kernels.cu; the workers' scans are replaced by plain loops
// with the same operation order as the kernels (sub, mul, add individually rounded; build with -ffp-contract=off).
//   extern "C" int ahc_emul(const double* rows, int N, int D, double* Z)   ->  0 ok, 5 NaN
#include "../../fluidaudio_b200/csrc/ahc_core.cuh"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

using namespace fa::ahc;

extern "C" int ahc_emul(const double *rows_in, int N, int D, double *Z) {
    if (N < 2) return 0;
    std::vector<double> rows((size_t)(2 * N - 1) * D);
    std::copy(rows_in, rows_in + (size_t)N * D, rows.begin());
    std::vector<int> node_weight(2 * N - 1, 1), node_of(N), slot_of(2 * N - 1), nn(N, 0);
    std::vector<int> worker_id(N);   // what each worker thread believes its slot holds
    std::vector<double> key(N, std::numeric_limits<double>::infinity());
    std::vector<uint16_t> at(N), where(N, 0);
    std::vector<unsigned> bits((2 * N - 1 + 31) / 32 + 1, 0xffffffffu);
    bool nan_seen = false;
    auto sq = [&](const double *a, const double *b) {
        double s = 0;
        for (int k = 0; k < D; ++k) {
            const double diff = a[k] - b[k];
            s = s + diff * diff;
        }
        if (s != s) nan_seen = true;
        return s;
    };
    // init NN (ahc_init_nn_kernel + ahc_init_reduce_kernel): lexicographic (d, j) minimum over j < i
    for (int i = 1; i < N; ++i) {
        double best = std::numeric_limits<double>::infinity();
        int arg = 0;
        for (int j = 0; j < i; ++j) {
            const double d = sq(&rows[(size_t)i * D], &rows[(size_t)j * D]);
            if (d < best) {
                best = d;
                arg = j;
            }
        }
        key[i] = best;
        nn[i] = arg;
    }
    if (nan_seen) return 5;
    for (int i = 0; i < N; ++i) node_of[i] = slot_of[i] = worker_id[i] = i;
    NnHeapT<uint16_t> heap{key.data(), at.data(), where.data(), 0};
    heap.build(N - 1, 1);
    LiveSet live{bits.data(), 2 * N - 1, 0};
    // worker-side scan: candidates are slots whose node id is >= 0 and < limit
    auto scan = [&](const double *v, int limit, double &d, int &id) {
        d = std::numeric_limits<double>::infinity();
        id = std::numeric_limits<int>::max();
        for (int s = 0; s < N; ++s) {
            const int nid = worker_id[s];
            if (nid < 0 || nid >= limit) continue;
            const double dist = sq(&rows[(size_t)nid * D], v);
            if (cand_less(dist, nid, d, id)) {
                d = dist;
                id = nid;
            }
        }
    };
    std::vector<int> ma(N - 1), mb(N - 1);
    std::vector<double> md(N - 1);
    for (int step = 0; step < N - 1; ++step) {
        const int fresh = N + step;
        int sa;
        for (;;) {
            sa = heap.top();
            if (!live.dead(nn[sa])) break;
            const int a_id = node_of[sa];
            double d;
            int id;
            scan(&rows[(size_t)a_id * D], a_id, d, id);
            if (nan_seen) return 5;
            nn[sa] = id;
            heap.raise_key(sa, d);
        }
        const int a = node_of[sa], b = nn[sa];
        live.drop(a);
        live.drop(b);
        ma[step] = a;
        mb[step] = b;
        md[step] = key[sa];
        if (step < N - 2) {
            const int sb = slot_of[b];
            node_of[sa] = fresh;
            node_of[sb] = -1;
            slot_of[fresh] = sa;
            // workers: build the centroid, relabel slots, scan
            const double wa = (double)node_weight[a], wb = (double)node_weight[b], den = wa + wb;
            double *v = &rows[(size_t)fresh * D];
            for (int k = 0; k < D; ++k) v[k] = (rows[(size_t)a * D + k] * wa + rows[(size_t)b * D + k] * wb) / den;
            node_weight[fresh] = (int)(wa + wb);
            for (int s = 0; s < N; ++s) {
                if (worker_id[s] == a) worker_id[s] = fresh;
                else if (worker_id[s] == b) worker_id[s] = -1;
            }
            double d;
            int id;
            scan(v, fresh, d, id);
            if (nan_seen) return 5;
            nn[sa] = id;
            if (b < live.head) heap.erase(slot_of[live.head]); else heap.erase(sb);
            heap.replace_key(sa, d);
        }
    }
    for (int s = 0; s < N - 1; ++s) {
        const int lo = std::min(ma[s], mb[s]), hi = std::max(ma[s], mb[s]);
        const double sz = (lo < N ? 1.0 : Z[(size_t)(lo - N) * 4 + 3]) + (hi < N ? 1.0 : Z[(size_t)(hi - N) * 4 + 3]);
        Z[(size_t)s * 4 + 0] = lo;
        Z[(size_t)s * 4 + 1] = hi;
        Z[(size_t)s * 4 + 2] = std::sqrt(md[s]);
        Z[(size_t)s * 4 + 3] = sz;
    }
    return 0;
}
