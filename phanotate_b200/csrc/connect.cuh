// The join of the reference's C extension, src/phanotate_connect.c:78-121 (`get_connected`; never
// called by phanotate.py, kept so that the extension has a drop-in too).
//
// Reference: add_edge(left, right) appends (key=left, value=right) to a `nodes_left` hash and (key=right, value=left)
// to a `nodes_right` hash (:44-95; uthash keeps insertion order and does not merge equal keys); get_connected walks
// ALL pairs, right entries outermost, and emits (right_i, left_j, 0) when
//     |right_i - left_j| <= 300   and   right_i != right_j   and   left_i != left_j          (:104-113)
// (the `min_distance` argument is parsed and ignored, :84-92).
//
// Here: edge i = (left[i], right[i]); item = (right entry i, chunk of 2048 left entries).  A block stages its chunk
// (left key + right value, 8 B per entry) in shared memory once and every thread walks it with broadcast reads; pass 1
// counts, an exclusive scan over the (i-major, chunk-minor) counters places every item's rows, pass 2 writes them: rows
// come out in exactly the reference's order without a sort.  Same O(n^2) work as the reference, spread over the chip.
#pragma once
#include "pipeline.cuh"

#define CN_CHUNK 2048
#define CN_BLOCK 256

struct ConnArgs {
    const i32* left;    // [n]
    const i32* right;   // [n]
    i32 n, nchunk;
    u64* cnt;           // [n * nchunk + 1] counters, then (after the scan) row offsets
    i32* out;           // [2 * total] rows (right_i, left_j)
};

PB_HD bool conn_pair(i32 ri, i32 li, i32 lj, i32 rj) {
    // 32-bit wrap-around like the compiled C `int` arithmetic (positions never get near the limits)
    i32 d = (i32)((u32)ri - (u32)lj);
    if (d < 0) d = (i32)(0u - (u32)d);
    return d <= 300 && ri != rj && li != lj;
}

// one item on the host build / the plain statement of an item's work
PB_HDN void conn_item(const ConnArgs& a, i64 item, bool fill) {
    const i32 i = (i32)(item / a.nchunk), ch = (i32)(item % a.nchunk);
    const i32 j0 = ch * CN_CHUNK, j1 = (j0 + CN_CHUNK < a.n) ? j0 + CN_CHUNK : a.n;
    const i32 ri = a.right[i], li = a.left[i];
    u64 k = fill ? a.cnt[item] : 0;
    for (i32 j = j0; j < j1; j++) {
        if (!conn_pair(ri, li, a.left[j], a.right[j])) continue;
        if (fill) {
            a.out[2 * k] = ri;
            a.out[2 * k + 1] = a.left[j];
        }
        k++;
    }
    if (!fill) a.cnt[item] = k;
}

#ifdef __CUDACC__
template <bool FILL>
__global__ void __launch_bounds__(CN_BLOCK) k_connect(const ConnArgs a) {
    __shared__ int2 S[CN_CHUNK];                       // (left key, right value) of the chunk
    const i32 ch = (i32)blockIdx.y;
    const i32 j0 = ch * CN_CHUNK, j1 = (j0 + CN_CHUNK < a.n) ? j0 + CN_CHUNK : a.n, m = j1 - j0;
    for (i32 t = threadIdx.x; t < m; t += CN_BLOCK) S[t] = make_int2(a.left[j0 + t], a.right[j0 + t]);
    __syncthreads();
    for (i32 i = (i32)(blockIdx.x * CN_BLOCK + threadIdx.x); i < a.n; i += (i32)(gridDim.x * CN_BLOCK)) {
        const i32 ri = a.right[i], li = a.left[i];
        const i64 item = (i64)i * a.nchunk + ch;
        u64 k = FILL ? a.cnt[item] : 0;
        if (FILL && a.cnt[item + 1] == k) continue;    // nothing to write for this item
#pragma unroll 8
        for (i32 t = 0; t < m; t++) {
            const int2 e = S[t];
            if (conn_pair(ri, li, e.x, e.y)) {
                if (FILL) *(int2*)(a.out + 2 * k) = make_int2(ri, e.x);
                k++;
            }
        }
        if (!FILL) a.cnt[item] = k;
    }
}
#endif
