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

// Heap over elements identified by SLOT (a slot holds one live node; a merged node inherits the slot of its first
// parent, so "rename old -> new node" is a key update in place).  key[] and where[] are indexed by slot, at[] by
// heap position.  The arrays live wherever the caller puts them: shared memory for the device master when they
// fit (Idx = uint16_t, 12 bytes per slot), global memory otherwise, plain host memory for the heapify.
template <typename Idx>
struct NnHeapT {
    double *key;
    Idx *at;
    Idx *where;
    int size;

    FA_HD double val(int pos) const { return key[at[pos]]; }
    FA_HD void swap_pos(int a, int b) {
        const Idx ea = at[a], eb = at[b];
        at[a] = eb;
        at[b] = ea;
        where[eb] = (Idx)a;
        where[ea] = (Idx)b;
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
    // identity layout over slots first..first+count-1, then Floyd heap construction
    FA_HD void build(int count, int first) {
        size = count;
        for (int i = 0; i < count; ++i) {
            at[i] = (Idx)(i + first);
            where[i + first] = (Idx)i;
        }
        for (int pos = size >> 1; pos > 0;) {
            --pos;
            sift_down(pos);
        }
    }
    FA_HD int top() const { return (int)at[0]; }
    FA_HD void raise_key(int slot, double v) {   // v >= old key
        key[slot] = v;
        sift_down((int)where[slot]);
    }
    FA_HD void lower_key(int slot, double v) {   // v <= old key
        key[slot] = v;
        sift_up((int)where[slot]);
    }
    // reference: binary_min_heap::replace(old, new, v) with old and new sharing a slot
    FA_HD void replace_key(int slot, double v) {
        if (v <= key[slot]) lower_key(slot, v); else raise_key(slot, v);
    }
    FA_HD void erase(int slot) {
        --size;
        const int pos = (int)where[slot];
        const Idx moved = at[size];
        where[moved] = (Idx)pos;
        at[pos] = moved;
        if (key[moved] <= key[slot]) sift_up(pos); else sift_down(pos);
    }
};

// Set of live node ids 0..count-1 as a bitmap; `head` is the smallest live id (the reference's
// doubly_linked_list::start, fastcluster_internal.hpp:299-350 — only start / is_inactive / remove are needed by
// the control plane, the ordered traversal is done by the scan kernels).
struct LiveSet {
    unsigned *bits;
    int count;
    int head;

    FA_HD void drop(int id) {
        bits[id >> 5] &= ~(1u << (id & 31));
        if (id == head) {
            int w = id >> 5;
            unsigned word = bits[w] & ~((2u << (id & 31)) - 1u);   // bits above id in the same word
            const int words = (count + 31) >> 5;
            while (word == 0u && ++w < words) word = bits[w];
            if (word == 0u) {
                head = count;
            } else {
                int b = 0;
                while (!((word >> b) & 1u)) ++b;
                head = (w << 5) + b;
            }
        }
    }
    FA_HD bool dead(int id) const { return ((bits[id >> 5] >> (id & 31)) & 1u) == 0u; }
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
