// Control-plane data structures of the centroid-linkage kernel, usable from host and device.
//
// The reference algorithm (Sources/FastClusterWrapper/fastcluster_internal.hpp:1625-1800,
// generic_linkage_vector_alternative<METHOD_VECTOR_CENTROID>) decides WHICH pair merges next through an indexed
// binary min-heap over per-node nearest-neighbour distances (:778-937) and an ascending list of live node ids
// (:299-350).  When several candidate pairs have exactly equal distance the winner depends on the heap's sift
// rules, so bit-exact dendrogram parity requires the same rules: strict '<' when sifting up; when sifting down,
// prefer the left child unless the right one is strictly smaller; removal moves the last element into the hole
// and sifts up iff its key is <= the removed key.  This file states those rules once, for both the host
// (heapify after the initial nearest-neighbour pass) and the device master thread (merge loop).
#pragma once

#include "fa_common.cuh"

namespace fa {
namespace ahc {

// Heap over elements identified by node id.  key[] is indexed by node id, at[] by heap position,
// where[] by node id.  All arrays live in caller-provided memory (global memory on the device).
struct NnHeap {
    double *key;
    int *at;
    int *where;
    int size;

    FA_HD double val(int pos) const { return key[at[pos]]; }
    FA_HD void swap_pos(int a, int b) {
        const int ea = at[a], eb = at[b];
        at[a] = eb;
        at[b] = ea;
        where[eb] = a;
        where[ea] = b;
    }
    FA_HD void sift_up(int pos) {
        while (pos > 0) {
            const int parent = (pos - 1) >> 1;
            if (!(val(pos) < val(parent))) break;
            swap_pos(pos, parent);
            pos = parent;
        }
    }
    FA_HD void sift_down(int pos) {
        for (;;) {
            int child = 2 * pos + 1;
            if (child >= size) break;
            const double here = val(pos);
            if (val(child) >= here) {
                ++child;
                if (child >= size || val(child) >= here) break;
            } else if (child + 1 < size && val(child + 1) < val(child)) {
                ++child;
            }
            swap_pos(pos, child);
            pos = child;
        }
    }
    // identity layout over ids first..first+count-1, then Floyd heap construction
    FA_HD void build(int count, int first) {
        size = count;
        for (int i = 0; i < count; ++i) {
            at[i] = i + first;
            where[i + first] = i;
        }
        for (int pos = size >> 1; pos > 0;) {
            --pos;
            sift_down(pos);
        }
    }
    FA_HD int top() const { return at[0]; }
    FA_HD void raise_key(int id, double v) {   // v >= old key
        key[id] = v;
        sift_down(where[id]);
    }
    FA_HD void lower_key(int id, double v) {   // v <= old key
        key[id] = v;
        sift_up(where[id]);
    }
    FA_HD void erase(int id) {
        --size;
        const int pos = where[id];
        const int moved = at[size];
        where[moved] = pos;
        at[pos] = moved;
        if (key[moved] <= key[id]) sift_up(pos); else sift_down(pos);
    }
    FA_HD void rename(int old_id, int new_id, double v) {
        const int pos = where[old_id];
        where[new_id] = pos;
        at[pos] = new_id;
        if (v <= key[old_id]) lower_key(new_id, v); else raise_key(new_id, v);
    }
};

// Ascending list of live node ids 0..count-1 with O(1) unlink; next[id] == 0 marks a dead id.
struct LiveList {
    int *next;
    int *prev;
    int head;

    FA_HD void build(int count) {
        head = 0;
        for (int i = 0; i < count; ++i) {
            prev[i + 1] = i;
            next[i] = i + 1;
        }
    }
    FA_HD void drop(int id) {
        if (id == head) {
            head = next[id];
        } else {
            const int p = prev[id], n = next[id];
            next[p] = n;
            prev[n] = p;
        }
        next[id] = 0;
    }
    FA_HD bool dead(int id) const { return next[id] == 0; }
};

// (distance, node id) candidates are ordered lexicographically: the scan "first strict minimum in ascending id
// order" of the reference (fastcluster_internal.hpp:1665-1668, 1724-1727, 1785-1788) is exactly this minimum.
struct Cand {
    double d;
    int id;
};
FA_HD bool cand_less(double d1, int id1, double d2, int id2) { return d1 < d2 || (d1 == d2 && id1 < id2); }

} // namespace ahc
} // namespace fa
