#include "assign_host.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace fa {
namespace assign {

namespace {
// Shortest-augmenting-path assignment with dual variables (row potentials `pr`, column potentials `pc`); columns and
// rows are 1-based inside, column 0 is the virtual start column.  Loop orders match the reference so that equal-cost
// optima are resolved identically.
struct Matcher {
    int n;
    const int64_t *cost;
    std::vector<int64_t> pr, pc, slack;
    std::vector<int32_t> row_of_col, prev_col;
    std::vector<char> visited;

    Matcher(const int64_t *c, int size)
        : n(size), cost(c), pr(size + 1, 0), pc(size + 1, 0), slack(size + 1), row_of_col(size + 1, 0),
          prev_col(size + 1, 0), visited(size + 1) {}

    void insert_row(int row) {
        const int64_t kInf = std::numeric_limits<int64_t>::max() / 4;
        row_of_col[0] = row;
        int col = 0;
        std::fill(slack.begin(), slack.end(), kInf);
        std::fill(visited.begin(), visited.end(), 0);
        do {
            visited[col] = 1;
            const int r = row_of_col[col];
            int64_t step = kInf;
            int next = 0;
            for (int j = 1; j <= n; ++j) {
                if (visited[j]) continue;
                const int64_t reduced = cost[(size_t)(r - 1) * n + (j - 1)] - pr[r] - pc[j];
                if (reduced < slack[j]) {
                    slack[j] = reduced;
                    prev_col[j] = col;
                }
                if (slack[j] < step) {
                    step = slack[j];
                    next = j;
                }
            }
            for (int j = 0; j <= n; ++j) {
                if (visited[j]) {
                    pr[row_of_col[j]] += step;
                    pc[j] -= step;
                } else {
                    slack[j] -= step;
                }
            }
            col = next;
        } while (row_of_col[col] != 0);
        do {   // flip the augmenting path
            const int back = prev_col[col];
            row_of_col[col] = row_of_col[back];
            col = back;
        } while (col != 0);
    }
};
} // namespace

void min_cost_matching(const int64_t *cost, int n, int32_t *out) {
    if (n <= 0) return;
    Matcher m(cost, n);
    for (int row = 1; row <= n; ++row) m.insert_row(row);
    std::fill(out, out + n, -1);
    for (int j = 1; j <= n; ++j)
        if (m.row_of_col[j] != 0) out[m.row_of_col[j] - 1] = j - 1;
}

void max_score_matching(const double *scores, int rows, int cols, int32_t *out) {
    if (rows <= 0) return;
    if (cols <= 0) {
        std::fill(out, out + rows, -1);
        return;
    }
    // finite range; non-finite scores rank one below the smallest finite score (HungarianAssignment.swift:76-79)
    bool seen = false;
    double hi = 0.0, lo = 0.0;
    const size_t total = (size_t)rows * cols;
    for (size_t i = 0; i < total; ++i) {
        const double s = scores[i];
        if (!std::isfinite(s)) continue;
        hi = seen ? std::max(hi, s) : s;
        lo = seen ? std::min(lo, s) : s;
        seen = true;
    }
    const double worst = lo - 1.0;
    const int n = std::max(rows, cols);
    std::vector<int64_t> cost((size_t)n * n, 0);   // dummy rows / columns share the constant cost 0
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const double s = scores[(size_t)r * cols + c];
            cost[(size_t)r * n + c] = (int64_t)std::round((hi - (std::isfinite(s) ? s : worst)) * 1e6);
        }
    std::vector<int32_t> full(n);
    min_cost_matching(cost.data(), n, full.data());
    for (int r = 0; r < rows; ++r) out[r] = full[r] < cols ? full[r] : -1;
}

void constrained_assign(const double *scores, long long N, int K, const int32_t *chunk, int32_t *out) {
    std::fill(out, out + N, -2);
    if (N <= 0) return;
    std::vector<long long> order((size_t)N);
    std::iota(order.begin(), order.end(), 0LL);
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return chunk[a] < chunk[b]; });
    std::vector<double> local;
    std::vector<int32_t> match;
    for (long long begin = 0; begin < N;) {
        long long end = begin;
        while (end < N && chunk[order[end]] == chunk[order[begin]]) ++end;
        const int rows = (int)(end - begin);
        local.assign((size_t)rows * std::max(K, 1), 0.0);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < K; ++c) local[(size_t)r * K + c] = scores[(size_t)order[begin + r] * K + c];
        match.assign(rows, -1);
        max_score_matching(local.data(), rows, K, match.data());
        for (int r = 0; r < rows; ++r) out[order[begin + r]] = match[r] >= 0 ? match[r] : -2;
        begin = end;
    }
}

void build_chunk_assignments(const int32_t *chunk, const int32_t *speaker, const int32_t *assignments, long long N,
                             int num_chunks, int num_speakers, int cluster_count, int32_t *matrix) {
    std::fill(matrix, matrix + (size_t)num_chunks * num_speakers, -2);
    for (long long i = 0; i < N; ++i) {
        const int c = chunk[i], s = speaker[i], a = assignments[i];
        if (c < 0 || c >= num_chunks || s < 0 || s >= num_speakers || a < 0 || a >= cluster_count) continue;
        matrix[(size_t)c * num_speakers + s] = a;
    }
}

} // namespace assign
} // namespace fa
