// C-ABI library of the B200-native PHANOTATE hot path (include/phanotate_b200.h).
//
// Product build:   nvcc -gencode arch=compute_100a,code=sm_100a  ->  libpb200.so  (needs a GPU)
// Test-only build: g++ -x c++ -DPB_HOSTSIM  ->  tests/native/pb200_hostsim.so : the same stage
//                  functions run as host loops so the stage logic can be unit-tested on a box
//                  without a GPU.  The Python package never loads that build.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <thread>
#include <algorithm>

#include "../../include/phanotate_b200.h"
#include "graph.cuh"
#include "chunk.cuh"
#include "trna.cuh"
#include "connect.cuh"
#include "fast.cuh"

static_assert(sizeof(pb200_dec) == sizeof(Dec), "Dec layout");
static_assert(sizeof(pb200_call) == sizeof(CallRec), "CallRec layout");
static_assert(sizeof(pb200_call24) == sizeof(Call24), "Call24 layout");
static_assert(sizeof(pb200_edge) == sizeof(EdgeRec), "EdgeRec layout");
static_assert(sizeof(pb200_orf) == sizeof(OrfRec), "OrfRec layout");
static_assert(sizeof(pb200_node) == sizeof(NodeRec), "NodeRec layout");
static_assert(sizeof(pb200_contig) == sizeof(ContigRec), "ContigRec layout");

#define NPHASE 14

#ifndef PB_HOSTSIM
// ================================================================================================
//                                         CUDA backend
// ================================================================================================
#include <cuda_runtime.h>
#define PB_BLOCK 128
#ifndef PB_SOLVE_MINB
#define PB_SOLVE_MINB 10
#endif
#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_);                         \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

#define PB_KERNEL(stage)                                                                        \
    __global__ void __launch_bounds__(PB_BLOCK) k_##stage(const Batch B, i64 n) {               \
        for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) stage(B, i); \
    }
PB_KERNEL(st_zero_tails)
PB_KERNEL(st_scan)
static_assert((int)FLAG_SCAN_REFERENCE == (int)PB200_SCAN_REFERENCE, "flag value");
PB_KERNEL(st_base_prefix)
PB_KERNEL(st_mark_starts)
PB_KERNEL(st_mark_stops)
PB_KERNEL(st_word_contig)
PB_KERNEL(st_count64)
PB_KERNEL(st_contig_offsets)
PB_KERNEL(st_node_pos)
PB_KERNEL(st_fill)
PB_KERNEL(st_orf_pstop)
PB_KERNEL(st_orf_lnx)
PB_KERNEL(st_orf_powA)
PB_KERNEL(st_orf_powF)
PB_KERNEL(st_orf_prepare)
PB_KERNEL(st_orf_finish)
PB_KERNEL(st_ov_pbar)
PB_KERNEL(st_ov_pow)
PB_KERNEL(st_len_scatter)
PB_KERNEL(st_ov_weight)
PB_KERNEL(st_contig_stats)
PB_KERNEL(st_gap_pow_int)
PB_KERNEL(st_gap_pow_real)
PB_KERNEL(st_gap_lut)
PB_KERNEL(st_node_attrs)
PB_KERNEL(st_ov_count)
PB_KERNEL(st_ov_fill)
PB_KERNEL(st_br_fill)
PB_KERNEL(st_backtrack)
PB_KERNEL(st_tie_fix)
PB_KERNEL(st_tie_link)
PB_KERNEL(st_gather_calls)
PB_KERNEL(st_call_orf)
PB_KERNEL(st_fast_tables)
PB_KERNEL(st_orf_fast)
PB_KERNEL(st_lit_calls)
PB_KERNEL(st_lit_rest)
PB_KERNEL(st_ov_fast)
PB_KERNEL(st_gap_fast)
PB_KERNEL(st_contig_lng)
PB_KERNEL(st_rbs_weights)
PB_KERNEL(st_edge_count)
PB_KERNEL(st_edge_fill)
PB_KERNEL(st_chunk_plan)
PB_KERNEL(st_trna_nodes)
PB_KERNEL(st_trna_tails)
PB_KERNEL(st_trna_count)
PB_KERNEL(st_trna_fill)
PB_KERNEL(st_chunk_ids)
PB_KERNEL(st_chunk_delta)
PB_KERNEL(st_lv_init)
PB_KERNEL(st_lv_node)
PB_KERNEL(st_lv_orf)
PB_KERNEL(st_lv_ov)
PB_KERNEL(st_lv_br)
PB_KERNEL(st_lv_check)
PB_KERNEL(st_lv_target)
PB_KERNEL(st_chunk_viol_count)
PB_KERNEL(st_chunk_retry_mark)
PB_KERNEL(st_pj_init)
PB_KERNEL(st_pj_round)
PB_KERNEL(st_pj_calls)

#include "scan_tile.cuh"
// one warp per contig; the 128-bit instantiation is kept small enough for 12 blocks per SM (the solve is a chain of
// dependent memory round trips per contig: throughput comes from the number of contigs in flight)
template <int NL>
__global__ void __launch_bounds__(PB_BLOCK, PB_SOLVE_MINB) k_solve(const Batch B, i32 nc) {
    const int lane = threadIdx.x & (NL - 1);
    const unsigned mask = NL == 32 ? 0xFFFFFFFFu : (((1u << (NL & 31)) - 1u) << ((threadIdx.x & 31) - lane));
    const i64 group = ((i64)blockIdx.x * blockDim.x + threadIdx.x) / NL;
    const i64 ngroups = ((i64)gridDim.x * blockDim.x) / NL;
    for (i64 c = group; c < nc; c += ngroups)
        if (!contig_is_wide(B, (int)c) && !contig_chunked(B, (int)c)) {
            // (contigs with tRNA nodes take the plain statement: it knows their explicit edge list, trna.cuh)
            if (NL == 32 && ((B.flags & PB200_SOLVE_PLAIN) || contig_has_trna(B, (int)c))) solve_contig_t<D128>(B, (int)c, lane, 32);
            else solve_contig_win<NL>(B, (int)c, lane, mask);
        }
}
// long contigs: one warp per chunk, distances only, into the chunk's private arrays (chunk.cuh)
__global__ void __launch_bounds__(PB_BLOCK) k_chunk_solve(const Batch B) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 id = warp; id < B.nch; id += nwarps) {
        const ChunkGeo g = chunk_geo(B, (i32)id);
        if (!chunk_active(B, g.c)) continue;
        const SolveRange R = chunk_range(B, g);
        if (B.flags & PB200_SOLVE_PLAIN) solve_contig_t<D128, true>(B, g.c, lane, 32, &R);
        else solve_contig_win<32, true>(B, g.c, lane, 0xFFFFFFFFu, &R);
    }
}
// warp-per-item stages of the chunked path: WHAT = 0 chunk_prefix (contig), 1 reach_chunk pass 0 (chunk), 2 reach_chunk_prefix
// (contig), 3 reach_chunk pass 2 (chunk)
template <int WHAT>
__global__ void __launch_bounds__(PB_BLOCK) k_chunk_warps(const Batch B, i64 n) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 i = warp; i < n; i += nwarps) {
        if (WHAT == 0) chunk_prefix(B, (int)i, lane, 32);
        else if (WHAT == 1) reach_chunk(B, (i32)i, 0, lane, 32);
        else if (WHAT == 2) {
            if (B.ch_cnt[i + 1] > B.ch_cnt[i]) reach_chunk_prefix(B, (int)i, lane, 32);
        } else reach_chunk(B, (i32)i, 2, lane, 32);
    }
}
// the one-warp sweep for the chunked contigs whose assembled distances failed a check
__global__ void __launch_bounds__(PB_BLOCK) k_solve_fallback(const Batch B, i32 nc) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 c = warp; c < nc; c += nwarps) {
        if (!contig_chunked(B, (int)c) || !B.cs[c].chunk_viol) continue;
        if (lane == 0) {
            B.cs[c].n_ties = 0;
            atomicAdd(B.lit_cnt + 4, 1u);
        }
        __syncwarp();
        solve_contig_win<32>(B, (int)c, lane, 0xFFFFFFFFu);
    }
}
__global__ void __launch_bounds__(PB_BLOCK) k_solve_wide(const Batch B, i32 nc) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 c = warp; c < nc; c += nwarps)
        if (contig_is_wide(B, (int)c) && !contig_is_huge(B, (int)c)) solve_contig_t<D256>(B, (int)c, lane, 32);
}
// contigs with an ORF weight beyond 256 bits: 2048-bit distances (a large local-memory frame per thread: its own kernel,
// launched only when a batch has such a contig)
__global__ void __launch_bounds__(32) k_solve_huge(const Batch B, i32 nc) {
    const int lane = threadIdx.x & 31;
    for (i64 c = blockIdx.x; c < nc; c += gridDim.x)
        if (contig_is_huge(B, (int)c)) solve_contig_t<DHuge>(B, (int)c, lane, 32);
}
// Overlap enumeration (st_ov_count / st_ov_fill) with the exit nodes compacted inside the warp: only exit nodes have
// work, and they are about every second node, so a warp takes 64 consecutive nodes and hands the k-th exit node among them
// to lane k (ballots + find-nth-set-bit) instead of leaving half of its lanes idle.
template <bool FILL>
__global__ void __launch_bounds__(PB_BLOCK) k_ov_exits(const Batch B) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 base = warp * 64; base < B.nn; base += nwarps * 64) {
        const i64 n0 = base + lane, n1 = base + 32 + lane;
        const bool x0 = n0 < B.nn && !kind_is_entry((int)(B.n_pk[n0] & 3));
        const bool x1 = n1 < B.nn && !kind_is_entry((int)(B.n_pk[n1] & 3));
        if (!FILL) {                                   // entry nodes have no overlap edges
            if (n0 < B.nn && !x0) B.ov_cnt[n0] = 0;
            if (n1 < B.nn && !x1) B.ov_cnt[n1] = 0;
        }
        const unsigned m0 = __ballot_sync(0xFFFFFFFFu, x0), m1 = __ballot_sync(0xFFFFFFFFu, x1);
        const int c0 = __popc(m0), total = c0 + __popc(m1);
        for (int k = lane; k < total; k += 32) {
            const int bit = k < c0 ? (int)__fns(m0, 0, k + 1) : 32 + (int)__fns(m1, 0, k - c0 + 1);
            overlaps_of(B, (i32)(base + bit), FILL);
        }
    }
}
// per-codon product + Orf.score, ORFs in length-sorted order; each thread keeps its six prepared
// factors in shared memory (18 x 16 B, stride = block size: conflict-free 128-bit loads)
__global__ void __launch_bounds__(PB_BLOCK) k_hold(const Batch B) {
    __shared__ U4 S[18 * PB_BLOCK];
    const int t = threadIdx.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + t; i < B.nlit; i += (i64)gridDim.x * blockDim.x) {
        const i32 sl = B.o_order[i];
        const U4* src = (const U4*)(B.o_hf + (i64)sl * 6);
#pragma unroll
        for (int j = 0; j < 18; j++) S[j * PB_BLOCK + t] = src[j];
        hold_run(B, sl, S, PB_BLOCK, t);
    }
}
__global__ void __launch_bounds__(PB_BLOCK) k_reach(const Batch B, i32 nc) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 c = warp; c < nc; c += nwarps) reach_contig(B, (int)c, lane, 32);
}
__global__ void k_pack_orfs(const Batch B, OrfRec* out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < B.no; i += (i64)gridDim.x * blockDim.x) pack_orf(B, i, out);
}
__global__ void k_pack_contigs(const Batch B, ContigRec* out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < B.nc; i += (i64)gridDim.x * blockDim.x) pack_contig(B, i, out);
}
__global__ void k_pack_nodes(const Batch B, NodeRec* out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < (i64)B.nn + 2 * B.nt; i += (i64)gridDim.x * blockDim.x) pack_node(B, i, out);
}
__global__ void k_bf_literal(const BFArgs a) { bf_literal(a); }

// 4-bit letters -> the lower-case letters the stages read (PB200_INPUT_PACKED4): 16 bases (8 bytes) per thread
__constant__ unsigned char k_nib2chr[16] = {'a', 'c', 'g', 't', 'n', 'r', 'y', 's', 'w', 'k', 'm', 'b', 'v', 'd', 'h', '?'};
__global__ void __launch_bounds__(256) k_unpack4(const unsigned char* __restrict__ packed, int skip, unsigned char* __restrict__ out, i64 nb) {
    // base g of the batch is nibble g + skip of `packed` (low nibble first)
    for (i64 g0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 16; g0 < nb; g0 += (i64)gridDim.x * blockDim.x * 16) {
        unsigned char o[16];
        const i64 n0 = g0 + skip;
        const i64 b0 = n0 >> 1;
        unsigned char in[9];
        const i64 nbytes = (nb + skip + 1) >> 1;
#pragma unroll
        for (int k = 0; k < 9; k++) in[k] = (b0 + k < nbytes) ? packed[b0 + k] : 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int nib = (int)(n0 & 1) + k;
            o[k] = k_nib2chr[(in[nib >> 1] >> ((nib & 1) * 4)) & 15];
        }
        if (g0 + 16 <= nb) {
            *(uint4*)(out + g0) = *(const uint4*)o;
        } else {
            for (int k = 0; k < 16 && g0 + k < nb; k++) out[g0 + k] = o[k];
        }
    }
}

// ---- exclusive prefix sums (three passes; T = u32 or u64 with two packed 32-bit counters)
#define SCAN_TILE 2048
template <typename T>
__global__ void __launch_bounds__(256) k_scan_tile_sums(const T* in, i64 n, T* sums) {
    __shared__ T sh[256];
    i64 base = (i64)blockIdx.x * SCAN_TILE;
    T s = 0;
    for (int k = threadIdx.x; k < SCAN_TILE; k += 256) {
        i64 i = base + k;
        if (i < n) s += in[i];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = sh[0];
}
template <typename T>
__global__ void __launch_bounds__(1024) k_scan_sums(T* sums, i64 nt, T* total) {
    __shared__ T sh[1024];
    __shared__ T carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (i64 base = 0; base < nt; base += 1024) {
        i64 i = base + threadIdx.x;
        T v = (i < nt) ? sums[i] : (T)0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            T t = ((int)threadIdx.x >= o) ? sh[threadIdx.x - o] : (T)0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        T incl = sh[threadIdx.x];
        if (i < nt) sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
template <typename T>
__global__ void __launch_bounds__(256) k_scan_apply(T* data, i64 n, const T* sums) {
    __shared__ T sh[256];
    i64 base = (i64)blockIdx.x * SCAN_TILE;
    const int per = SCAN_TILE / 256;
    T loc[SCAN_TILE / 256];
    T s = 0;
    i64 i0 = base + (i64)threadIdx.x * per;
    for (int k = 0; k < per; k++) {
        i64 i = i0 + k;
        T v = (i < n) ? data[i] : (T)0;
        loc[k] = s;
        s += v;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        T t = ((int)threadIdx.x >= o) ? sh[threadIdx.x - o] : (T)0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    T off = sums[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < per; k++) {
        i64 i = i0 + k;
        if (i < n) data[i] = off + loc[k];
    }
}

struct DevBuf {
    char* p = nullptr;
    size_t cap = 0, used = 0;
};
struct StageTime {
    const char* name;
    cudaEvent_t a, b;
};
struct pb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr, side_ev = nullptr;
    std::string err;
    DevBuf ph[NPHASE];
    DevBuf in_seq, in_off, in_pack, scratch, scratch2, conn, conn_out;
    Batch B;
    bool have = false;
    std::vector<StageTime> times;
    std::vector<cudaEvent_t> evpool;
    size_t evused = 0;
    int launches = 0;
    int sm_count = 148;
    int contig_base = 0;
    bool scan_attr_set = false;
    std::vector<i32> t_contig, t_start, t_stop, t_first;   // tRNA hits for the next runs (pb200_set_trnas)
    bool ch_default = true;      // geometry never set by the caller: a small batch may shorten it (driver.inc)
    int ch_core = 256, ch_warm = 768, ch_margin = 64, ch_long = 4096;
    cudaEvent_t run_a = nullptr, run_b = nullptr, sync_ev = nullptr;
    void* comm = nullptr;        // CommState (comm.inc) once pb200_comm_init ran
    cudaStream_t copy_stream = nullptr;   // pb200_upload_async: the stream several contexts queue their copies on, in order
    cudaEvent_t upload_ev = nullptr;      // ... this context's letters have arrived
    cudaEvent_t unpack_ev = nullptr;      // the 4-bit letters of the current batch have been expanded: in_pack is free again
    const uint8_t* pf_ptr = nullptr;      // pb200_prefetch_async: the host buffer whose 4-bit letters already sit in in_pack
    size_t pf_bytes = 0;
    int pf_skip = 0;
    cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
};
// wait for the context's stream.  PB200_BLOCKING_SYNC=1 (environment, read at pb200_create) makes the host thread sleep on
// a blocking-sync event instead of spinning: for boxes with fewer cores than (ranks x lanes) host threads
static cudaError_t ctx_sync(pb200_ctx* ctx) {
    if (!ctx->sync_ev) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->sync_ev, ctx->stream);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->sync_ev);
}
static int buf_ensure(pb200_ctx* ctx, DevBuf& b, size_t bytes) {
    b.used = 0;
    if (bytes <= b.cap) return 0;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    CK(cudaMalloc((void**)&b.p, want));
    b.cap = want;
    return 0;
}
static void* buf_take(DevBuf& b, size_t bytes) {
    size_t a = (b.used + 255) & ~(size_t)255;
    if (a + bytes > b.cap) return nullptr;
    b.used = a + bytes;
    return b.p + a;
}
static cudaEvent_t ev_get(pb200_ctx* ctx) {
    if (ctx->evused == ctx->evpool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->evpool.push_back(e);
    }
    return ctx->evpool[ctx->evused++];
}
static int grid_for(pb200_ctx* ctx, i64 n, int block) {
    i64 g = (n + block - 1) / block;
    i64 cap = (i64)ctx->sm_count * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
template <typename T>
static int dev_scan(pb200_ctx* ctx, T* data, i64 n) {   // exclusive, total -> data[n]
    i64 nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nt < 1) nt = 1;
    if (buf_ensure(ctx, ctx->scratch, (size_t)(nt + 1) * sizeof(T))) return -1;
    T* sums = (T*)buf_take(ctx->scratch, (size_t)(nt + 1) * sizeof(T));
    k_scan_tile_sums<T><<<(int)nt, 256, 0, ctx->stream>>>(data, n, sums);
    k_scan_sums<T><<<1, 1024, 0, ctx->stream>>>(sums, nt, data + n);
    k_scan_apply<T><<<(int)nt, 256, 0, ctx->stream>>>(data, n, sums);
    ctx->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

#define PB_PHASE(k, bytes)                                             \
    do {                                                               \
        if (buf_ensure(ctx, ctx->ph[k], (size_t)(bytes))) return -1;   \
    } while (0)
#define PB_ALLOC(k, T, count) ((T*)buf_take(ctx->ph[k], (size_t)(count) * sizeof(T)))
#define PB_ZERO(ptr, bytes) CK(cudaMemsetAsync((ptr), 0, (bytes), ctx->stream))
#define PB_FETCH(dst, src, bytes)                                                                \
    do {                                                                                         \
        CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream));         \
        CK(ctx_sync(ctx));                                                  \
    } while (0)
#define PB_UPLOAD(dst, src, bytes) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
#define PB_RUN(stage, n)                                                                         \
    do {                                                                                         \
        i64 n_ = (i64)(n);                                                                       \
        if (n_ > 0) {                                                                            \
            StageTime t_;                                                                        \
            t_.name = #stage;                                                                    \
            t_.a = ev_get(ctx);                                                                  \
            t_.b = ev_get(ctx);                                                                  \
            cudaEventRecord(t_.a, ctx->stream);                                                  \
            k_##stage<<<grid_for(ctx, n_, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B, n_);         \
            cudaEventRecord(t_.b, ctx->stream);                                                  \
            ctx->times.push_back(t_);                                                            \
            ctx->launches++;                                                                     \
            CK(cudaGetLastError());                                                              \
        }                                                                                        \
    } while (0)
#define PB_RUN_SCAN()                                                                            \
    do {                                                                                         \
        if (B.flags & PB200_SCAN_REFERENCE) PB_RUN(st_scan, (B.nb + SCAN_STRIP - 1) / SCAN_STRIP); \
        else {                                                                                   \
            StageTime t_;                                                                        \
            t_.name = "scan_tiles";                                                              \
            t_.a = ev_get(ctx);                                                                  \
            t_.b = ev_get(ctx);                                                                  \
            const i64 ntiles_ = (B.nb + ST_T - 1) / ST_T;                                        \
            /* grid = 12 waves of the resident block count (4 per SM): contiguous tile ranges per block, yet fine enough \
               for the block scheduler to even out SM speed differences (one wave measured 5 % slower) */ \
            const i64 slots_ = (i64)ctx->sm_count * 4 * 12;                                      \
            const int tpb_ = (int)((ntiles_ + slots_ - 1) / slots_);                             \
            if (!ctx->scan_attr_set) {                                                           \
                cudaFuncSetAttribute(k_scan_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem)); \
                cudaFuncSetAttribute(k_scan_tiles<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
                cudaFuncSetAttribute(k_scan_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem)); \
                cudaFuncSetAttribute(k_scan_tiles<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
                ctx->scan_attr_set = true;                                                       \
            }                                                                                    \
            /* bulk TMA copies need a 16-byte aligned source with slack behind the last base: the library's own  \
               input buffer has both, a caller-owned device buffer (PB200_INPUT_DEVICE) takes the plain loads */ \
            const int tma_ = (!(B.flags & PB200_INPUT_DEVICE) && (((size_t)B.seq) & 15) == 0) ? 1 : 0; \
            cudaEventRecord(t_.a, ctx->stream);                                                  \
            if (tma_) k_scan_tiles<true><<<(int)((ntiles_ + tpb_ - 1) / tpb_), ST_NT, sizeof(ScanSmem), ctx->stream>>>(B, ntiles_, tpb_); \
            else k_scan_tiles<false><<<(int)((ntiles_ + tpb_ - 1) / tpb_), ST_NT, sizeof(ScanSmem), ctx->stream>>>(B, ntiles_, tpb_); \
            cudaEventRecord(t_.b, ctx->stream);                                                  \
            ctx->times.push_back(t_);                                                            \
            ctx->launches++;                                                                     \
            CK(cudaGetLastError());                                                              \
        }                                                                                        \
    } while (0)
#define PB_LAUNCH_CHUNKS(stream_)                                                                \
    do {                                                                                         \
            /* one warp per chunk: out of shared memory when a chunk's tables fit (chunk.cuh), else out of HBM/L2 */ \
            const int stride_ = B.ch_warm + B.ch_core + B.ch_margin;                             \
            const int ecap_ = stride_ + stride_ / 4;                                             \
            const size_t smem_ = chunk_smem_bytes(stride_, ecap_);                               \
            /* (a lone warp sweeping out of shared memory is bound by instruction latency, ~0.4 us per visit against ~1 us \
               out of L2: it wins while all chunks are resident at once -- one or a few genomes -- and loses to the L2 \
               kernel's 9 warps per SM once the chunks need several rounds of the 3-9 blocks per SM that fit) */ \
            i64 fit_ = (i64)(227 * 1024 / (smem_ + 1024));                                       \
            if (fit_ > 32) fit_ = 32;                                                            \
            const char* force_ = getenv("PB200_CHUNK_KERNEL");                                   \
            const bool smem_ok_ = smem_ <= 200 * 1024 && !(B.flags & PB200_SOLVE_PLAIN);         \
            const bool use_smem_ = force_ ? (smem_ok_ && force_[0] == 's') : (smem_ok_ && (i64)B.nch <= fit_ * ctx->sm_count); \
            if (use_smem_) {                                                                     \
                cudaFuncSetAttribute(k_chunk_solve_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_); \
                i64 gb_ = B.nch;                                                                 \
                if (gb_ > (i64)ctx->sm_count * 32) gb_ = (i64)ctx->sm_count * 32;                \
                k_chunk_solve_smem<<<(int)gb_, 32, smem_, (stream_)>>>(B, stride_, ecap_);    \
            } else {                                                                             \
                k_chunk_solve<<<grid_for(ctx, (i64)B.nch * 32, PB_BLOCK), PB_BLOCK, 0, (stream_)>>>(B); \
            }                                                                                    \
            ctx->launches++;                                                                     \
    } while (0)
#define PB_RUN_CHUNKS()                                                                          \
    do {                                                                                         \
        StageTime t_;                                                                            \
        t_.name = "chunk_solve_retry";                                                           \
        t_.a = ev_get(ctx);                                                                      \
        t_.b = ev_get(ctx);                                                                      \
        cudaEventRecord(t_.a, ctx->stream);                                                      \
        PB_LAUNCH_CHUNKS(ctx->stream);                                                           \
        cudaEventRecord(t_.b, ctx->stream);                                                      \
        ctx->times.push_back(t_);                                                                \
        CK(cudaGetLastError());                                                                  \
    } while (0)
#define PB_RUN_SOLVE(nc_)                                                                        \
    do {                                                                                         \
        StageTime t_;                                                                            \
        t_.name = "solve";                                                                       \
        t_.a = ev_get(ctx);                                                                      \
        t_.b = ev_get(ctx);                                                                      \
        cudaEventRecord(t_.a, ctx->stream);                                                      \
        /* the few contigs that need 256-bit distances run beside the others on a second stream */ \
        cudaEventRecord(ctx->fork_ev, ctx->stream);                                              \
        cudaStreamWaitEvent(ctx->stream2, ctx->fork_ev, 0);                                      \
        k_solve_wide<<<grid_for(ctx, (i64)(nc_) * 32, PB_BLOCK), PB_BLOCK, 0, ctx->stream2>>>(B, (nc_)); \
        if (B.n_huge > 0) {                                                                      \
            k_solve_huge<<<(nc_) < 1024 ? (nc_) : 1024, 32, 0, ctx->stream2>>>(B, (nc_));       \
            ctx->launches++;                                                                     \
        }                                                                                        \
        if (B.nch > 0) PB_LAUNCH_CHUNKS(ctx->stream2);                                            \
        cudaEventRecord(ctx->join_ev, ctx->stream2);                                             \
        /* PB200_SOLVE_HALF=1 (environment): two contigs per warp, 16 lanes each -- twice the sweeps in flight at the   \
           same register cost.  Measured SLOWER on the bench workload (7.6 against 6.8 ms: the two halves' divergent    \
           control flow costs more issue slots than the extra latency hiding returns), so it is opt-in */              \
        if (!(B.flags & PB200_SOLVE_PLAIN) && getenv("PB200_SOLVE_HALF"))                                            \
            k_solve<16><<<grid_for(ctx, (i64)(nc_) * 16, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B, (nc_));             \
        else                                                                                                         \
            k_solve<32><<<grid_for(ctx, (i64)(nc_) * 32, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B, (nc_));             \
        cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0);                                       \
        ctx->launches++;                                                                         \
        cudaEventRecord(t_.b, ctx->stream);                                                      \
        ctx->times.push_back(t_);                                                                \
        ctx->launches++;                                                                         \
        CK(cudaGetLastError());                                                                  \
    } while (0)
#define PB_RUN_CHUNK_WARPS(what_, name_, n_)                                                     \
    do {                                                                                         \
        StageTime t_;                                                                            \
        t_.name = name_;                                                                         \
        t_.a = ev_get(ctx);                                                                      \
        t_.b = ev_get(ctx);                                                                      \
        cudaEventRecord(t_.a, ctx->stream);                                                      \
        k_chunk_warps<what_><<<grid_for(ctx, (i64)(n_) * 32, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B, (i64)(n_)); \
        cudaEventRecord(t_.b, ctx->stream);                                                      \
        ctx->times.push_back(t_);                                                                \
        ctx->launches++;                                                                         \
        CK(cudaGetLastError());                                                                  \
    } while (0)
#define PB_RUN_FALLBACK(nc_)                                                                     \
    do {                                                                                         \
        StageTime t_;                                                                            \
        t_.name = "solve_fallback";                                                              \
        t_.a = ev_get(ctx);                                                                      \
        t_.b = ev_get(ctx);                                                                      \
        cudaEventRecord(t_.a, ctx->stream);                                                      \
        k_solve_fallback<<<grid_for(ctx, (i64)(nc_) * 32, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B, (nc_)); \
        cudaEventRecord(t_.b, ctx->stream);                                                      \
        ctx->times.push_back(t_);                                                                \
        ctx->launches++;                                                                         \
        CK(cudaGetLastError());                                                                  \
    } while (0)
// A run of stages that nothing on the main stream needs for a while goes to the second stream (with its own scan
// scratch): BEGIN forks, END hands the main stream back, JOIN makes the main stream wait for the side work.
#define PB_SIDE_BEGIN()                                                  \
    do {                                                                 \
        cudaEventRecord(ctx->fork_ev, ctx->stream);                      \
        cudaStreamWaitEvent(ctx->stream2, ctx->fork_ev, 0);              \
        std::swap(ctx->stream, ctx->stream2);                            \
        std::swap(ctx->scratch, ctx->scratch2);                          \
    } while (0)
#define PB_SIDE_END()                                                    \
    do {                                                                 \
        cudaEventRecord(ctx->side_ev, ctx->stream);                      \
        std::swap(ctx->stream, ctx->stream2);                            \
        std::swap(ctx->scratch, ctx->scratch2);                          \
    } while (0)
#define PB_SIDE_JOIN() cudaStreamWaitEvent(ctx->stream, ctx->side_ev, 0)
#define PB_SIDE_RESUME()                                                 \
    do { /* back to the side stream without making it wait for the main stream */ \
        std::swap(ctx->stream, ctx->stream2);                            \
        std::swap(ctx->scratch, ctx->scratch2);                          \
    } while (0)
#define PB_RUN_OV(fill_)                                                                         \
    do {                                                                                         \
        if (B.nn > 0) {                                                                          \
            StageTime t_;                                                                        \
            t_.name = (fill_) ? "st_ov_fill" : "st_ov_count";                                    \
            t_.a = ev_get(ctx);                                                                  \
            t_.b = ev_get(ctx);                                                                  \
            cudaEventRecord(t_.a, ctx->stream);                                                  \
            if (fill_) k_ov_exits<true><<<grid_for(ctx, ((i64)B.nn + 1) / 2, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B);  \
            else k_ov_exits<false><<<grid_for(ctx, ((i64)B.nn + 1) / 2, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B);       \
            cudaEventRecord(t_.b, ctx->stream);                                                  \
            ctx->times.push_back(t_);                                                            \
            ctx->launches++;                                                                     \
            CK(cudaGetLastError());                                                              \
        }                                                                                        \
    } while (0)
#define PB_RUN_REACH(nc_)                                                                        \
    do {                                                                                         \
        StageTime t_;                                                                            \
        t_.name = "reach";                                                                       \
        t_.a = ev_get(ctx);                                                                      \
        t_.b = ev_get(ctx);                                                                      \
        cudaEventRecord(t_.a, ctx->stream);                                                      \
        k_reach<<<grid_for(ctx, (i64)(nc_) * 32, PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B, (nc_)); \
        cudaEventRecord(t_.b, ctx->stream);                                                      \
        ctx->times.push_back(t_);                                                                \
        ctx->launches++;                                                                         \
        CK(cudaGetLastError());                                                                  \
    } while (0)
#define PB_RUN_HOLD(no_)                                                                         \
    do {                                                                                         \
        if ((no_) > 0) {                                                                         \
            StageTime t_;                                                                        \
            t_.name = "hold";                                                                    \
            t_.a = ev_get(ctx);                                                                  \
            t_.b = ev_get(ctx);                                                                  \
            cudaEventRecord(t_.a, ctx->stream);                                                  \
            k_hold<<<grid_for(ctx, (i64)(no_), PB_BLOCK), PB_BLOCK, 0, ctx->stream>>>(B);        \
            cudaEventRecord(t_.b, ctx->stream);                                                  \
            ctx->times.push_back(t_);                                                            \
            ctx->launches++;                                                                     \
            CK(cudaGetLastError());                                                              \
        }                                                                                        \
    } while (0)
#define PB_SCAN64(ptr, n)                          \
    do {                                           \
        if (dev_scan<u64>(ctx, (ptr), (n))) return -1; \
    } while (0)
#define PB_SCAN32(ptr, n)                          \
    do {                                           \
        if (dev_scan<u32>(ctx, (ptr), (n))) return -1; \
    } while (0)
#define PB_TO_HOST(dst, src, bytes)                                                              \
    do {                                                                                         \
        CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream));         \
        CK(ctx_sync(ctx));                                                  \
    } while (0)

#else
// ================================================================================================
//                              host-sim backend (unit tests only)
// ================================================================================================
#define CK(x) (x)
struct DevBuf {
    char* p = nullptr;
    size_t cap = 0, used = 0;
};
struct pb200_ctx {
    int device = 0;
    int contig_base = 0;
    int ch_core = 256, ch_warm = 768, ch_margin = 64, ch_long = 4096;
    bool ch_default = true;
    std::vector<i32> t_contig, t_start, t_stop, t_first;
    std::string err;
    DevBuf ph[NPHASE];
    DevBuf in_seq, in_off, in_pack, scratch, conn, conn_out;
    Batch B;
    bool have = false;
    int launches = 0;
};
static int buf_ensure(pb200_ctx*, DevBuf& b, size_t bytes) {
    b.used = 0;
    if (bytes <= b.cap) return 0;
    free(b.p);
    b.cap = bytes + bytes / 8 + 4096;
    b.p = (char*)malloc(b.cap);
    return b.p ? 0 : -1;
}
static void* buf_take(DevBuf& b, size_t bytes) {
    size_t a = (b.used + 255) & ~(size_t)255;
    if (a + bytes > b.cap) return nullptr;
    b.used = a + bytes;
    return b.p + a;
}
template <typename T>
static int dev_scan(pb200_ctx*, T* data, i64 n) {
    T s = 0;
    for (i64 i = 0; i < n; i++) {
        T v = data[i];
        data[i] = s;
        s += v;
    }
    data[n] = s;
    return 0;
}
#define PB_PHASE(k, bytes)                                             \
    do {                                                               \
        if (buf_ensure(ctx, ctx->ph[k], (size_t)(bytes))) return -1;   \
    } while (0)
#define PB_ALLOC(k, T, count) ((T*)buf_take(ctx->ph[k], (size_t)(count) * sizeof(T)))
#define PB_ZERO(ptr, bytes) memset((ptr), 0, (bytes))
#define PB_FETCH(dst, src, bytes) memcpy((dst), (src), (bytes))
#define PB_UPLOAD(dst, src, bytes) memcpy((dst), (src), (bytes))
#define PB_TO_HOST(dst, src, bytes) memcpy((dst), (src), (bytes))
#define PB_RUN(stage, n)                                  \
    do {                                                  \
        i64 n_ = (i64)(n);                                \
        for (i64 i_ = 0; i_ < n_; i_++) stage(B, i_);     \
        ctx->launches++;                                  \
    } while (0)
#define PB_RUN_SCAN() PB_RUN(st_scan, (B.nb + SCAN_STRIP - 1) / SCAN_STRIP)
#define PB_RUN_SOLVE(nc_)                                              \
    do {                                                               \
        for (i32 c_ = 0; c_ < (nc_); c_++) solve_contig(B, c_, 0, 1);  \
        for (i32 id_ = 0; id_ < B.nch; id_++) chunk_solve(B, id_, 0, 1); \
        ctx->launches++;                                               \
    } while (0)
#define PB_RUN_CHUNKS()                                                  \
    do {                                                                 \
        for (i32 id_ = 0; id_ < B.nch; id_++) chunk_solve(B, id_, 0, 1); \
        ctx->launches++;                                                 \
    } while (0)
#define PB_RUN_FALLBACK(nc_)                                             \
    do {                                                                 \
        for (i32 c_ = 0; c_ < (nc_); c_++) solve_fallback(B, c_, 0, 1);  \
        ctx->launches++;                                                 \
    } while (0)
#define PB_RUN_CHUNK_WARPS(what_, name_, n_)                                                      \
    do {                                                                                          \
        for (i64 i_ = 0; i_ < (i64)(n_); i_++) {                                                  \
            if ((what_) == 0) chunk_prefix(B, (int)i_, 0, 1);                                     \
            else if ((what_) == 1) reach_chunk(B, (i32)i_, 0, 0, 1);                              \
            else if ((what_) == 2) {                                                              \
                if (B.ch_cnt[i_ + 1] > B.ch_cnt[i_]) reach_chunk_prefix(B, (int)i_, 0, 1);        \
            } else reach_chunk(B, (i32)i_, 2, 0, 1);                                              \
        }                                                                                         \
        ctx->launches++;                                                                          \
    } while (0)
#define PB_SIDE_BEGIN()
#define PB_SIDE_END()
#define PB_SIDE_JOIN()
#define PB_SIDE_RESUME()
#define PB_RUN_OV(fill_)                                 \
    do {                                                 \
        if (fill_) PB_RUN(st_ov_fill, B.nn);             \
        else PB_RUN(st_ov_count, B.nn);                  \
    } while (0)
#define PB_RUN_REACH(nc_)                                                  \
    do {                                                                   \
        for (i32 c_ = 0; c_ < (nc_); c_++) reach_contig(B, c_, 0, 1);      \
        ctx->launches++;                                                   \
    } while (0)
#define PB_RUN_HOLD(no_)                                                   \
    do {                                                                   \
        for (i64 i_ = 0; i_ < (i64)(no_); i_++) {                          \
            const i32 oi_ = B.o_order[i_];                                 \
            hold_run(B, oi_, (const U4*)(B.o_hf + (i64)oi_ * 6), 1, 0);    \
        }                                                                  \
        ctx->launches++;                                                   \
    } while (0)
#define PB_SCAN64(ptr, n) dev_scan<u64>(ctx, (ptr), (n))
#define PB_SCAN32(ptr, n) dev_scan<u32>(ctx, (ptr), (n))
#endif

// ================================================================================================
//                                   backend-independent host code
// ================================================================================================
static int codon_code(const char* s) {        // 'acgt' codon -> 0..63, -1 if not plain acgt
    int v = 0;
    for (int i = 0; i < 3; i++) {
        int c;
        switch (s[i] | 0x20) {
            case 'a': c = 0; break;
            case 'c': c = 1; break;
            case 'g': c = 2; break;
            case 't': c = 3; break;
            default: return -1;
        }
        v = v * 4 + c;
    }
    return s[3] == 0 ? v : -1;
}
static int revcomp_code(int v) {
    int a = v >> 4, b = (v >> 2) & 3, c = v & 3;
    return (3 - c) * 16 + (3 - b) * 4 + (3 - a);
}
// codon classes with the reference's elif priority (functions.py:198-215)
static int make_params(pb200_ctx* ctx, const pb200_params* in, Params* P) {
    memset(P, 0, sizeof(*P));
    if (in->n_start < 1 || in->n_start > 8 || in->n_stop < 0 || in->n_stop > 8) {
        ctx->err = "between 1 and 8 start codons and at most 8 stop codons are supported";
        return -2;
    }
    if (in->min_orf_len < 9) {
        ctx->err = "min_orf_len must be >= 9";
        return -2;
    }
    bool is_start[64] = {false}, is_stop[64] = {false};
    int sw[64];
    for (int i = 0; i < 64; i++) sw[i] = -1;
    for (int k = 0; k < in->n_start; k++) {
        int v = codon_code(in->start_codon[k]);
        if (v < 0) {
            ctx->err = "start codons must be three letters of acgt";
            return -2;
        }
        is_start[v] = true;
        sw[v] = k;                           // a repeated key keeps the last weight, like the dict in file_handling.py:58-60
        memcpy(&P->startw[k], &in->start_weight[k], sizeof(Dec));
    }
    for (int k = 0; k < in->n_stop; k++) {
        int v = codon_code(in->stop_codon[k]);
        if (v < 0) {
            ctx->err = "stop codons must be three letters of acgt";
            return -2;
        }
        is_stop[v] = true;
    }
    for (int v = 0; v < 64; v++) {
        int rc = revcomp_code(v);
        int cls = CLS_NONE;
        if (is_start[v]) cls = CLS_S;
        else if (is_start[rc]) cls = CLS_s;
        else if (is_stop[v]) cls = CLS_T;
        else if (is_stop[rc]) cls = CLS_t;
        P->codon_cls[v] = (u8)cls;
        P->rev_start[v] = is_start[rc] ? 1 : 0;
        P->sw_fwd[v] = (signed char)sw[v];
        P->sw_rev[v] = (signed char)sw[rc];
    }
    P->min_orf_len = in->min_orf_len;
    return 0;
}

// The literal chain (hold.cuh) over the first nlit entries of B.lit_ids (or over every ORF when B.lit_all):
// the six Decimal factors per ORF, the per-codon product replayed multiplication by multiplication, Orf.score().
#define ALN(x) (((size_t)(x) + 255) & ~(size_t)255)
// second half of phase 3: size and fill the bridge tables (on whichever stream the context currently points at)
static int bridge_tables(pb200_ctx* ctx) {
    Batch& B = ctx->B;
    {
        u32 tot;
        PB_FETCH(&tot, B.br_cnt + B.nn, 4);
        B.nbr = (i32)tot;
    }
    {
        const size_t nbr = (size_t)B.nbr + 1;
        PB_PHASE(3, 2 * ALN(nbr * 4) + ALN(nbr * sizeof(Dec)) + ALN(nbr * sizeof(WInt)) + 1024);
        B.br_src = PB_ALLOC(3, i32, nbr);
        B.br_dst = PB_ALLOC(3, i32, nbr);
        B.br_wint = PB_ALLOC(3, WInt, nbr);
    }
    if (B.nbr > 0) PB_RUN(st_br_fill, B.nn);
    return 0;
}
static int literal_chain(pb200_ctx* ctx, i32 nlit) {
    Batch& B = ctx->B;
    B.nlit = nlit;
    if (nlit <= 0) return 0;
    const size_t m = (size_t)nlit + 1;
    PB_PHASE(8, ALN(m * 6 * sizeof(Dec)) + ALN(m * 6 * sizeof(HoldFac)) + ALN(m * sizeof(Dec)) + ALN(m * sizeof(SFx)) +
                    ALN(m * 3 * sizeof(Dec)) + ALN(m * 3 * sizeof(SFx)) + ALN(m * 2) + ALN(m * 4) + 2 * ALN((HOLD_BINS + 1) * 4) + 4096);
    B.o_fac = PB_ALLOC(8, Dec, m * 6);
    B.o_hf = PB_ALLOC(8, HoldFac, m * 6);
    B.o_hold = PB_ALLOC(8, Dec, m);
    B.o_lnx = PB_ALLOC(8, SFx, m);
    B.o_A = PB_ALLOC(8, Dec, m * 3);
    B.o_lnA = PB_ALLOC(8, SFx, m * 3);
    B.o_bin = PB_ALLOC(8, unsigned short, m);
    B.o_order = PB_ALLOC(8, i32, m);
    B.len_hist = PB_ALLOC(8, u32, HOLD_BINS + 1);
    B.len_cursor = PB_ALLOC(8, u32, HOLD_BINS + 1);
    PB_ZERO(B.len_hist, (HOLD_BINS + 1) * 4);
    PB_ZERO(B.len_cursor, (HOLD_BINS + 1) * 4);
    PB_RUN(st_orf_pstop, nlit);
    PB_RUN(st_orf_lnx, nlit);
    PB_RUN(st_orf_powA, (i64)nlit * 3);
    PB_RUN(st_orf_powF, (i64)nlit * 6);
    PB_RUN(st_orf_prepare, (i64)nlit * 6);
    PB_SCAN32(B.len_hist, HOLD_BINS);
    PB_RUN(st_len_scatter, nlit);
    PB_RUN_HOLD(nlit);
    PB_RUN(st_orf_finish, nlit);
    return 0;
}
#undef ALN
// score_gap for every length -2..300 in Decimal arithmetic: gap_same / gap_diff and their integers
static int gap_tables(pb200_ctx* ctx) {
    Batch& B = ctx->B;
    if (B.gap_dec) return 0;
    const i32 nc = B.nc;
    PB_RUN(st_contig_lng, nc);
    PB_RUN(st_gap_pow_int, (i64)nc * 101);
    PB_RUN(st_gap_pow_real, (i64)nc * 202);
    PB_RUN(st_gap_lut, (i64)nc * GAPN);
    B.gap_dec = 1;
    return 0;
}
// score_overlap replayed in Decimal arithmetic over the first n entries of B.ovlit_ids (or every edge when B.ov_all)
static int overlap_chain(pb200_ctx* ctx, i32 n) {
    Batch& B = ctx->B;
    B.novlit = n;
    if (n <= 0) return 0;
    PB_PHASE(10, ((size_t)n + 1) * sizeof(Dec) + 1024);
    B.ov_w = PB_ALLOC(10, Dec, (size_t)n + 1);
    PB_RUN(st_ov_pbar, n);
    PB_RUN(st_ov_pow, n);
    PB_RUN(st_ov_weight, n);
    return 0;
}
static int ensure_literal_overlaps(pb200_ctx* ctx) {
    Batch& B = ctx->B;
    if (B.ov_all || B.nov < 1) return 0;
    B.ov_all = 1;
    return overlap_chain(ctx, B.nov);
}
// Decimal weight of every ORF that still lacks one (after a certified run; pb200_get_orfs, pb200_build_edges)
static int ensure_literal_orfs(pb200_ctx* ctx) {
    Batch& B = ctx->B;
    if (B.lit_all || B.lit_done || B.no < 1) return 0;
    PB_ZERO(B.lit_cnt + 1, 4);
    PB_RUN(st_lit_rest, B.no);
    u32 cnt;
    PB_FETCH(&cnt, B.lit_cnt + 1, 4);
    if (literal_chain(ctx, (i32)cnt)) return -1;
    B.lit_done = 1;
    return 0;
}

// PB200_INPUT_PACKED4: two bases per byte (low nibble first), base g of the batch = nibble g + skip of `packed`.
// Copies the packed letters to the device and expands them into the context's letter buffer (lower case, '?' for the
// code of a letter outside the IUPAC alphabet -> ERR_CHAR like the letter itself would have given).
static const char PB_NIB2CHR[17] = "acgtnryswkmbvdh?";
static int stage_packed4(pb200_ctx* ctx, const uint8_t* packed, int skip, const int64_t* offsets, int32_t n_contigs) {
    const i64 nb = offsets[n_contigs];
    const size_t pbytes = (size_t)((nb + skip + 1) >> 1);
    if (buf_ensure(ctx, ctx->in_seq, (size_t)nb + 64)) return -1;
    if (buf_ensure(ctx, ctx->in_off, (size_t)(n_contigs + 1) * 8)) return -1;
    if (buf_ensure(ctx, ctx->in_pack, pbytes + 64)) return -1;
#ifndef PB_HOSTSIM
    ctx->pf_ptr = nullptr;
    CK(cudaMemcpyAsync(ctx->in_pack.p, packed, pbytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->in_off.p, offsets, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    i64 g = (nb / 16 + 255) / 256;
    if (g > (i64)ctx->sm_count * 16) g = (i64)ctx->sm_count * 16;
    if (g < 1) g = 1;
    k_unpack4<<<(int)g, 256, 0, ctx->stream>>>((const unsigned char*)ctx->in_pack.p, skip, (unsigned char*)ctx->in_seq.p, nb);
    CK(cudaGetLastError());
#else
    memcpy(ctx->in_off.p, offsets, (size_t)(n_contigs + 1) * 8);
    for (i64 k = 0; k < nb; k++) {
        const i64 nib = k + skip;
        ctx->in_seq.p[k] = PB_NIB2CHR[(packed[nib >> 1] >> ((nib & 1) * 4)) & 15];
    }
#endif
    return 0;
}

static int run_pipeline(pb200_ctx* ctx) {
    Batch& B = ctx->B;
#include "driver.inc"
    return 0;
}

extern "C" {

int pb200_create(int device, pb200_ctx** out) {
    if (!out) return -2;
    pb200_ctx* ctx = new pb200_ctx();
    ctx->device = device;
#ifndef PB_HOSTSIM
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreate(&ctx->stream);
    if (e == cudaSuccess) e = cudaStreamCreate(&ctx->stream2);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->side_ev, cudaEventDisableTiming);
    {
        const char* bs = getenv("PB200_BLOCKING_SYNC");
        if (e == cudaSuccess && bs && bs[0] == '1') e = cudaEventCreateWithFlags(&ctx->sync_ev, cudaEventDisableTiming | cudaEventBlockingSync);
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "phanotate_b200: no usable CUDA device %d: %s\n", device, cudaGetErrorString(e));
        delete ctx;
        *out = nullptr;
        return -1;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
#endif
    if (const char* e2 = getenv("PB200_CHUNK")) {
        int a, b, c, d;
        if (sscanf(e2, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) pb200_set_chunking(ctx, a, b, c, d);
    }
    *out = ctx;
    return 0;
}

int pb200_comm_destroy(pb200_ctx* ctx);
void pb200_destroy(pb200_ctx* ctx) {
    if (!ctx) return;
    pb200_comm_destroy(ctx);
#ifndef PB_HOSTSIM
    cudaSetDevice(ctx->device);
    for (int k = 0; k < NPHASE; k++) cudaFree(ctx->ph[k].p);
    cudaFree(ctx->in_seq.p);
    cudaFree(ctx->in_off.p);
    cudaFree(ctx->in_pack.p);
    cudaFree(ctx->scratch.p);
    cudaFree(ctx->scratch2.p);
    cudaFree(ctx->conn.p);
    cudaFree(ctx->conn_out.p);
    for (auto e : ctx->evpool) cudaEventDestroy(e);
    if (ctx->run_a) cudaEventDestroy(ctx->run_a);
    if (ctx->run_b) cudaEventDestroy(ctx->run_b);
    if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
    if (ctx->upload_ev) cudaEventDestroy(ctx->upload_ev);
    if (ctx->unpack_ev) cudaEventDestroy(ctx->unpack_ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int k = 0; k < 4; k++)
        if (ctx->marks[k]) cudaEventDestroy(ctx->marks[k]);
    if (ctx->side_ev) cudaEventDestroy(ctx->side_ev);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
#else
    for (int k = 0; k < NPHASE; k++) free(ctx->ph[k].p);
    free(ctx->in_seq.p);
    free(ctx->in_off.p);
    free(ctx->in_pack.p);
    free(ctx->scratch.p);
    free(ctx->conn.p);
    free(ctx->conn_out.p);
#endif
    delete ctx;
}

const char* pb200_last_error(pb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pb200_run(pb200_ctx* ctx, const uint8_t* bases, const int64_t* offsets, int32_t n_contigs,
              const pb200_params* params, uint32_t flags) {
    if (!ctx || !offsets || !params || n_contigs < 1) {
        if (ctx) ctx->err = "bad arguments";
        return -2;
    }
    if ((flags & PB200_INPUT_PACKED4) && (flags & PB200_INPUT_DEVICE)) {
        ctx->err = "PB200_INPUT_PACKED4 takes a host buffer (not with PB200_INPUT_DEVICE)";
        return -2;
    }
    ctx->have = false;
    Batch& B = ctx->B;
    memset(&B, 0, sizeof(B));
    int rc = make_params(ctx, params, &B.P);
    if (rc) return rc;
    B.nc = n_contigs;
    B.flags = (i32)flags;
    B.contig_base = ctx->contig_base;
    B.nt = (i32)ctx->t_contig.size();
    if (B.nt > 0) {
        ctx->t_first.assign((size_t)n_contigs + 1, 0);
        for (i32 k = 0; k < B.nt; k++) {
            const i32 c = ctx->t_contig[k];
            if (c < 0 || c >= n_contigs || (k > 0 && c < ctx->t_contig[k - 1])) {
                ctx->err = "pb200_set_trnas: contig indices must be sorted and inside the batch";
                return -2;
            }
            ctx->t_first[(size_t)c + 1]++;
        }
        for (i32 c = 0; c < n_contigs; c++) ctx->t_first[(size_t)c + 1] += ctx->t_first[c];
    }
    B.ch_core = ctx->ch_core;
    B.ch_warm = ctx->ch_warm;
    B.ch_margin = ctx->ch_margin;
    B.ch_long = ctx->ch_long;
    // a small batch leaves most of the GPU idle: then every contig longer than one chunk's sweep is worth cutting
    if (n_contigs < 1024 && B.ch_long > B.ch_warm + B.ch_core + B.ch_margin) B.ch_long = B.ch_warm + B.ch_core + B.ch_margin;
#ifndef PB_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    ctx->times.clear();
    ctx->evused = 0;
    ctx->launches = 0;
    if (!ctx->run_a) {
        CK(cudaEventCreate(&ctx->run_a));
        CK(cudaEventCreate(&ctx->run_b));
    }
    CK(cudaEventRecord(ctx->run_a, ctx->stream));
    if (flags & PB200_INPUT_DEVICE) {
        i64 last;
        CK(cudaMemcpyAsync(&last, offsets + n_contigs, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx_sync(ctx));
        B.nb = last;
        B.seq = bases;
        B.coff = offsets;
    } else if (flags & PB200_REUSE_INPUT) {
        B.nb = offsets[n_contigs];
        if (!ctx->in_seq.p || ctx->in_seq.cap < (size_t)B.nb || !ctx->in_off.p) {
            ctx->err = "PB200_REUSE_INPUT without a resident batch";
            return -2;
        }
        B.seq = (const u8*)ctx->in_seq.p;
        B.coff = (const i64*)ctx->in_off.p;
    } else if (flags & PB200_INPUT_PACKED4) {
        B.nb = offsets[n_contigs];
        if (B.nb < 1 || !bases) {
            ctx->err = "empty batch";
            return -2;
        }
        if (stage_packed4(ctx, bases, 0, offsets, n_contigs)) return -1;
        B.seq = (const u8*)ctx->in_seq.p;
        B.coff = (const i64*)ctx->in_off.p;
    } else {
        B.nb = offsets[n_contigs];
        if (B.nb < 1 || !bases) {
            ctx->err = "empty batch";
            return -2;
        }
        if (buf_ensure(ctx, ctx->in_seq, (size_t)B.nb + 64)) return -1;
        if (buf_ensure(ctx, ctx->in_off, (size_t)(n_contigs + 1) * 8)) return -1;
        CK(cudaMemcpyAsync(ctx->in_seq.p, bases, (size_t)B.nb, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->in_off.p, offsets, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        B.seq = (const u8*)ctx->in_seq.p;
        B.coff = (const i64*)ctx->in_off.p;
    }
#else
    ctx->launches = 0;
    B.nb = offsets[n_contigs];
    if (flags & PB200_REUSE_INPUT) {
        B.seq = (const u8*)ctx->in_seq.p;
        B.coff = (const i64*)ctx->in_off.p;
    } else if (flags & PB200_INPUT_PACKED4) {
        if (stage_packed4(ctx, bases, 0, offsets, n_contigs)) return -1;
        B.seq = (const u8*)ctx->in_seq.p;
        B.coff = (const i64*)ctx->in_off.p;
    } else {
        B.seq = bases;
        B.coff = offsets;
    }
#endif
    if (B.nb < 1) {
        ctx->err = "empty batch";
        return -2;
    }
    if (B.nb >= ((i64)1 << 32)) {
        ctx->err = "batch larger than 2^32 bases";
        return -2;
    }
    rc = run_pipeline(ctx);
    if (rc) return rc;
#ifndef PB_HOSTSIM
    CK(cudaEventRecord(ctx->run_b, ctx->stream));
    CK(ctx_sync(ctx));
#endif
    ctx->have = true;
    return 0;
}

// Copy a batch into the context's device buffers without running it (then pb200_run(..., PB200_REUSE_INPUT)): lets a
// caller with several contexts order the host->device copies one after the other while kernels of other contexts run.
int pb200_upload(pb200_ctx* ctx, const uint8_t* bases, const int64_t* offsets, int32_t n_contigs) {
    if (!ctx || !bases || !offsets || n_contigs < 1) return -2;
#ifndef PB_HOSTSIM
    const i64 nb = offsets[n_contigs];
    if (nb < 1) {
        ctx->err = "empty batch";
        return -2;
    }
    CK(cudaSetDevice(ctx->device));
    if (buf_ensure(ctx, ctx->in_seq, (size_t)nb + 64)) return -1;
    if (buf_ensure(ctx, ctx->in_off, (size_t)(n_contigs + 1) * 8)) return -1;
    CK(cudaMemcpyAsync(ctx->in_seq.p, bases, (size_t)nb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->in_off.p, offsets, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx_sync(ctx));
#else
    const i64 nb = offsets[n_contigs];
    if (buf_ensure(ctx, ctx->in_seq, (size_t)nb + 64)) return -1;
    if (buf_ensure(ctx, ctx->in_off, (size_t)(n_contigs + 1) * 8)) return -1;
    memcpy(ctx->in_seq.p, bases, (size_t)nb);
    memcpy(ctx->in_off.p, offsets, (size_t)(n_contigs + 1) * 8);
#endif
    return 0;
}

// Geometry of the chunked solve of long contigs (chunk.cuh), in nodes (one node per ~28 bp): contigs with more than
// `long_nodes` nodes are cut into chunks of `core` nodes, each swept by its own warp from `warm` nodes upstream to `margin`
// nodes downstream.  Any geometry gives the same results (every node's distance is checked; a contig that fails is solved
// again by one sweep); it only moves the time.  Environment PB200_CHUNK="core,warm,margin,long" sets it at pb200_create.
int pb200_set_chunking(pb200_ctx* ctx, int32_t core, int32_t warm, int32_t margin, int32_t long_nodes) {
    if (!ctx || core < 1 || warm < 0 || margin < 0 || long_nodes < 0) return -2;
    ctx->ch_default = false;
    ctx->ch_core = core;
    ctx->ch_warm = warm;
    ctx->ch_margin = margin;
    ctx->ch_long = long_nodes;
    return 0;
}

// tRNA hits for the following runs (functions.py:457-509): hit k lies on contig contig[k] (index inside the batch, sorted
// ascending) from start[k] to stop[k], 1-based as aragorn / tRNAscan-SE report them, start > stop on the reverse strand --
// the [start, stop] pairs add_trnas collects, in its order.  n = 0 clears the list.
int pb200_set_trnas(pb200_ctx* ctx, const int32_t* contig, const int32_t* start, const int32_t* stop, int32_t n) {
    if (!ctx || n < 0 || (n > 0 && (!contig || !start || !stop))) return -2;
    ctx->t_contig.assign(contig, contig + n);
    ctx->t_start.assign(start, start + n);
    ctx->t_stop.assign(stop, stop + n);
    return 0;
}

// pb200_upload for 4-bit letters (see pb200_pack4): base g of the batch = nibble g + skip of `packed` (skip = 0 or 1: a
// group of contigs cut out of a larger packed batch may start in the middle of a byte)
int pb200_upload_packed4(pb200_ctx* ctx, const uint8_t* packed, int32_t skip, const int64_t* offsets, int32_t n_contigs) {
    if (!ctx || !packed || !offsets || n_contigs < 1 || skip < 0 || skip > 1) return -2;
    if (offsets[n_contigs] < 1) {
        ctx->err = "empty batch";
        return -2;
    }
#ifndef PB_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    if (stage_packed4(ctx, packed, skip, offsets, n_contigs)) return -1;
#ifndef PB_HOSTSIM
    CK(ctx_sync(ctx));
#endif
    return 0;
}

// pb200_upload without blocking the host: the copy is queued on `via`'s copy stream (several contexts that pass the same
// `via` get their copies one after the other, in call order: the first group's letters arrive first and its kernels start
// while the others are still on the link), and ctx's own stream waits for it on the device.  skip < 0: one byte per base;
// skip = 0 / 1: 4-bit letters (pb200_upload_packed4).  The host buffers must stay untouched until the run that uses them
// has finished.  Then pb200_run(ctx, ..., PB200_REUSE_INPUT).
int pb200_upload_async(pb200_ctx* ctx, pb200_ctx* via, const uint8_t* data, int32_t skip, const int64_t* offsets, int32_t n_contigs) {
    if (!ctx || !via || !data || !offsets || n_contigs < 1 || skip > 1) return -2;
    const i64 nb = offsets[n_contigs];
    if (nb < 1) {
        ctx->err = "empty batch";
        return -2;
    }
#ifndef PB_HOSTSIM
    if (ctx->device != via->device) {
        ctx->err = "pb200_upload_async: contexts on different devices";
        return -2;
    }
    CK(cudaSetDevice(ctx->device));
    if (!via->copy_stream) CK(cudaStreamCreateWithFlags(&via->copy_stream, cudaStreamNonBlocking));
    if (!ctx->upload_ev) CK(cudaEventCreateWithFlags(&ctx->upload_ev, cudaEventDisableTiming));
    if (buf_ensure(ctx, ctx->in_seq, (size_t)nb + 64)) return -1;
    if (buf_ensure(ctx, ctx->in_off, (size_t)(n_contigs + 1) * 8)) return -1;
    // (the previous run of this context has finished: pb200_run returns synchronised, so its buffers are free)
    if (skip < 0) {
        CK(cudaMemcpyAsync(ctx->in_seq.p, data, (size_t)nb, cudaMemcpyHostToDevice, via->copy_stream));
    } else {
        const size_t pbytes = (size_t)((nb + skip + 1) >> 1);
        // (letters that pb200_prefetch_async already brought in during the previous run are not copied again)
        const bool have = ctx->pf_ptr == data && ctx->pf_bytes == pbytes && ctx->pf_skip == skip;
        ctx->pf_ptr = nullptr;
        if (!have) {
            if (buf_ensure(ctx, ctx->in_pack, pbytes + 64)) return -1;
            CK(cudaMemcpyAsync(ctx->in_pack.p, data, pbytes, cudaMemcpyHostToDevice, via->copy_stream));
        }
    }
    CK(cudaMemcpyAsync(ctx->in_off.p, offsets, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, via->copy_stream));
    CK(cudaEventRecord(ctx->upload_ev, via->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->upload_ev, 0));
    if (skip >= 0) {
        i64 g = (nb / 16 + 255) / 256;
        if (g > (i64)ctx->sm_count * 16) g = (i64)ctx->sm_count * 16;
        if (g < 1) g = 1;
        k_unpack4<<<(int)g, 256, 0, ctx->stream>>>((const unsigned char*)ctx->in_pack.p, skip, (unsigned char*)ctx->in_seq.p, nb);
        CK(cudaGetLastError());
        if (!ctx->unpack_ev) CK(cudaEventCreateWithFlags(&ctx->unpack_ev, cudaEventDisableTiming));
        CK(cudaEventRecord(ctx->unpack_ev, ctx->stream));
    }
    return 0;
#else
    (void)via;
    if (skip < 0) return pb200_upload(ctx, data, offsets, n_contigs);
    return stage_packed4(ctx, data, skip, offsets, n_contigs);
#endif
}
// Double buffering across batches: queue the copy of the NEXT batch's 4-bit letters (n_bases of them from nibble `skip` of
// `data`) into the context's packed-letter buffer while the current run is still going -- that buffer is free as soon as
// the current batch has been expanded (k_unpack4, the first kernel of the run), and the copy stream waits for exactly
// that.  The next pb200_upload_async with the same `data`, `skip` and size then only sends the offsets.  Call it after
// the pb200_upload_async of the current batch, with the same `via`; `data` must stay untouched until that next run.
int pb200_prefetch_async(pb200_ctx* ctx, pb200_ctx* via, const uint8_t* data, int32_t skip, int64_t n_bases) {
    if (!ctx || !via || !data || n_bases < 1 || skip < 0 || skip > 1) return -2;
#ifndef PB_HOSTSIM
    if (ctx->device != via->device) {
        ctx->err = "pb200_prefetch_async: contexts on different devices";
        return -2;
    }
    CK(cudaSetDevice(ctx->device));
    if (!via->copy_stream) CK(cudaStreamCreateWithFlags(&via->copy_stream, cudaStreamNonBlocking));
    const size_t pbytes = (size_t)((n_bases + skip + 1) >> 1);
    if (buf_ensure(ctx, ctx->in_pack, pbytes + 64)) return -1;
    if (ctx->unpack_ev) CK(cudaStreamWaitEvent(via->copy_stream, ctx->unpack_ev, 0));
    CK(cudaMemcpyAsync(ctx->in_pack.p, data, pbytes, cudaMemcpyHostToDevice, via->copy_stream));
    ctx->pf_ptr = data;
    ctx->pf_bytes = pbytes;
    ctx->pf_skip = skip;
#else
    (void)via;   // (host build: nothing to overlap; the next upload copies as usual)
#endif
    return 0;
}

int pb200_set_contig_base(pb200_ctx* ctx, int32_t base) {
    if (!ctx) return -2;
    ctx->contig_base = base;
    return 0;
}

// Device-side stopwatch on the context's stream: pb200_mark(ctx, k) records event k (0..3) behind everything queued so far,
// pb200_elapsed_ms(ctx, a, b) = time between two recorded marks (waits for b).  For timing a region of several runs and
// gathers with CUDA events on the launching stream.
int pb200_mark(pb200_ctx* ctx, int32_t k) {
    if (!ctx || k < 0 || k > 3) return -2;
#ifndef PB_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    if (!ctx->marks[k]) CK(cudaEventCreate(&ctx->marks[k]));
    CK(cudaEventRecord(ctx->marks[k], ctx->stream));
#endif
    return 0;
}
float pb200_elapsed_ms(pb200_ctx* ctx, int32_t a, int32_t b) {
#ifndef PB_HOSTSIM
    float v = -1.f;
    if (!ctx || a < 0 || a > 3 || b < 0 || b > 3 || !ctx->marks[a] || !ctx->marks[b]) return -1.f;
    if (cudaEventSynchronize(ctx->marks[b]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&v, ctx->marks[a], ctx->marks[b]) != cudaSuccess) return -1.f;
    return v;
#else
    (void)ctx;
    (void)a;
    (void)b;
    return -1.f;
#endif
}

// ... between mark a of one context and mark b of another on the same device (several contexts working on one batch)
float pb200_elapsed_between_ms(pb200_ctx* from, int32_t a, pb200_ctx* to, int32_t b) {
#ifndef PB_HOSTSIM
    float v = -1.f;
    if (!from || !to || a < 0 || a > 3 || b < 0 || b > 3 || !from->marks[a] || !to->marks[b] || from->device != to->device) return -1.f;
    if (cudaEventSynchronize(to->marks[b]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&v, from->marks[a], to->marks[b]) != cudaSuccess) return -1.f;
    return v;
#else
    (void)from;
    (void)a;
    (void)to;
    (void)b;
    return -1.f;
#endif
}

float pb200_last_run_ms(pb200_ctx* ctx) {
#ifndef PB_HOSTSIM
    float v = -1.f;
    if (ctx && ctx->have && cudaEventElapsedTime(&v, ctx->run_a, ctx->run_b) == cudaSuccess) return v;
#else
    (void)ctx;
#endif
    return -1.f;
}

const pb200_call* pb200_device_calls(pb200_ctx* ctx) {
    return (ctx && ctx->have) ? (const pb200_call*)ctx->B.calls : nullptr;
}

int pb200_pin_host(void* ptr, size_t bytes) {
#ifndef PB_HOSTSIM
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : -1;
#else
    (void)ptr;
    (void)bytes;
    return 0;
#endif
}
int pb200_unpin_host(void* ptr) {
#ifndef PB_HOSTSIM
    return cudaHostUnregister(ptr) == cudaSuccess ? 0 : -1;
#else
    (void)ptr;
    return 0;
#endif
}
int pb200_struct_sizes(int32_t out[8]) {
    out[0] = (int32_t)sizeof(pb200_dec);
    out[1] = (int32_t)sizeof(pb200_params);
    out[2] = (int32_t)sizeof(pb200_call);
    out[3] = (int32_t)sizeof(pb200_orf);
    out[4] = (int32_t)sizeof(pb200_node);
    out[5] = (int32_t)sizeof(pb200_edge);
    out[6] = (int32_t)sizeof(pb200_contig);
    out[7] = 0;
    return 0;
}

int pb200_sizes(pb200_ctx* ctx, int64_t out[8]) {
    if (!ctx || !ctx->have) return -2;
    const Batch& B = ctx->B;
    out[0] = B.nc;
    out[1] = B.nb;
    out[2] = B.nn;
    out[3] = B.no;
    out[4] = B.nov;
    out[5] = B.nbr;
    out[6] = B.ncalls;
    out[7] = B.nedges;
    return 0;
}

int pb200_stats(pb200_ctx* ctx, int64_t out[8]) {
    if (!ctx || !ctx->have) return -2;
    const Batch& B = ctx->B;
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[0] = B.lit_all ? B.no : B.n_lit_pre;
    out[1] = B.lit_all ? 0 : B.n_lit_post;
    out[2] = (B.flags & PB200_LITERAL) ? B.nov : B.n_ovlit;
    out[5] = B.nt;                                    // tRNA hits of the run
    out[7] = B.ch_round;                              // 1: some long contig needed the second attempt of the chunked solve
    out[6] = B.n_huge;                                // ORF weights beyond 256 bits (their contigs solved at 2048 bits)
    out[3] = B.nch;                                   // chunks the long contigs were solved in
    if (B.nch > 0) {                                  // long contigs that failed the check and were solved by one sweep
        u32 fb = 0;
        PB_TO_HOST(&fb, B.lit_cnt + 4, 4);
        out[4] = fb;
    }
    return 0;
}

int pb200_get_orf_int_weights(pb200_ctx* ctx, uint32_t* out) {
    if (!ctx || !ctx->have) return -2;
    if (ctx->B.no > 0) PB_TO_HOST(out, ctx->B.o_wint, (size_t)ctx->B.no * sizeof(WInt));
    return 0;
}

int pb200_get_gap_int_weights(pb200_ctx* ctx, int64_t* same, int64_t* diff) {
    if (!ctx || !ctx->have) return -2;
    const size_t n = (size_t)ctx->B.nc * GAPN;
    PB_TO_HOST(same, ctx->B.gapi_same, n * 8);
    PB_TO_HOST(diff, ctx->B.gapi_diff, n * 8);
    return 0;
}

int pb200_get_overlap_int_weights(pb200_ctx* ctx, int64_t* out) {
    if (!ctx || !ctx->have) return -2;
    if (ctx->B.nov > 0) PB_TO_HOST(out, ctx->B.ov_w64, (size_t)ctx->B.nov * 8);
    return 0;
}

int pb200_get_calls(pb200_ctx* ctx, pb200_call* out) {
    if (!ctx || !ctx->have) return -2;
    if (ctx->B.ncalls > 0) PB_TO_HOST(out, ctx->B.calls, (size_t)ctx->B.ncalls * sizeof(CallRec));
    return 0;
}
// call rows without the Decimal weight (pb200_call24), n of them from src to dst (both on the device)
PB_HD void call24_of(const CallRec* src, Call24* dst, i64 i) {
    Call24 r;
    r.contig = src[i].contig;
    r.left = src[i].left;
    r.right = src[i].right;
    r.strand = src[i].strand;
    r.score = src[i].score;
    dst[i] = r;
}
#ifndef PB_HOSTSIM
__global__ void __launch_bounds__(256) k_calls24(const CallRec* src, Call24* dst, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) call24_of(src, dst, i);
}
static int calls24_launch(pb200_ctx* ctx, const CallRec* src, Call24* dst, i64 n, cudaStream_t st) {
    if (n < 1) return 0;
    k_calls24<<<grid_for(ctx, n, 256), 256, 0, st>>>(src, dst, n);
    CK(cudaGetLastError());
    return 0;
}
#endif
int pb200_get_calls24(pb200_ctx* ctx, pb200_call24* out) {
    if (!ctx || !ctx->have) return -2;
    Batch& B = ctx->B;
    if (B.ncalls < 1) return 0;
    PB_PHASE(13, (size_t)B.ncalls * sizeof(Call24) + 1024);
    Call24* tmp = PB_ALLOC(13, Call24, B.ncalls);
#ifndef PB_HOSTSIM
    if (calls24_launch(ctx, B.calls, tmp, B.ncalls, ctx->stream)) return -1;
#else
    for (i64 i = 0; i < B.ncalls; i++) call24_of(B.calls, tmp, i);
#endif
    PB_TO_HOST(out, tmp, (size_t)B.ncalls * sizeof(Call24));
    return 0;
}

int pb200_get_contigs(pb200_ctx* ctx, pb200_contig* out) {
    if (!ctx || !ctx->have) return -2;
    Batch& B = ctx->B;
    PB_PHASE(11, (size_t)B.nc * sizeof(ContigRec) + 1024);
    ContigRec* tmp = PB_ALLOC(11, ContigRec, B.nc);
#ifndef PB_HOSTSIM
    k_pack_contigs<<<grid_for(ctx, B.nc, 128), 128, 0, ctx->stream>>>(B, tmp);
    CK(cudaGetLastError());
#else
    for (i64 i = 0; i < B.nc; i++) pack_contig(B, i, tmp);
#endif
    PB_TO_HOST(out, tmp, (size_t)B.nc * sizeof(ContigRec));
    return 0;
}

int pb200_get_orfs(pb200_ctx* ctx, pb200_orf* out) {
    if (!ctx || !ctx->have) return -2;
    Batch& B = ctx->B;
    if (B.no < 1) return 0;
    if (ensure_literal_orfs(ctx)) return -1;
    PB_PHASE(6, (size_t)B.no * sizeof(OrfRec) + 1024);
    OrfRec* tmp = PB_ALLOC(6, OrfRec, B.no);
#ifndef PB_HOSTSIM
    k_pack_orfs<<<grid_for(ctx, B.no, 128), 128, 0, ctx->stream>>>(B, tmp);
    CK(cudaGetLastError());
#else
    for (i64 i = 0; i < B.no; i++) pack_orf(B, i, tmp);
#endif
    PB_TO_HOST(out, tmp, (size_t)B.no * sizeof(OrfRec));
    return 0;
}

// Orf.hold (orfs.py:84, functions.py:286-298): the per-codon product before Orf.score() inverts it, for every ORF in
// pb200_get_orfs order.  Only a PB200_LITERAL run keeps it for every ORF (a certified run never forms it).
int pb200_get_orf_holds(pb200_ctx* ctx, pb200_dec* out) {
    if (!ctx || !ctx->have) return -2;
    Batch& B = ctx->B;
    if (!B.lit_all || !(B.flags & PB200_LITERAL)) {
        ctx->err = "pb200_get_orf_holds needs a PB200_LITERAL run";
        return -2;
    }
    if (B.no > 0) PB_TO_HOST(out, B.o_hold, (size_t)B.no * sizeof(Dec));
    return 0;
}

int pb200_get_nodes(pb200_ctx* ctx, pb200_node* out) {
    if (!ctx || !ctx->have) return -2;
    Batch& B = ctx->B;
    if (B.nn < 1) return 0;
    const i64 nall = (i64)B.nn + 2 * B.nt;            // (the tRNA nodes follow the regular ones)
    PB_PHASE(6, (size_t)nall * sizeof(NodeRec) + 1024);
    NodeRec* tmp = PB_ALLOC(6, NodeRec, nall);
#ifndef PB_HOSTSIM
    k_pack_nodes<<<grid_for(ctx, nall, 128), 128, 0, ctx->stream>>>(B, tmp);
    CK(cudaGetLastError());
#else
    for (i64 i = 0; i < nall; i++) pack_node(B, i, tmp);
#endif
    PB_TO_HOST(out, tmp, (size_t)nall * sizeof(NodeRec));
    return 0;
}

int pb200_build_edges(pb200_ctx* ctx) {
    if (!ctx || !ctx->have) return -2;
    Batch& B = ctx->B;
    B.nedges = 0;
    if (B.nn < 1) return 0;
    if (ensure_literal_orfs(ctx)) return -1;
    if (ensure_literal_overlaps(ctx)) return -1;
    if (gap_tables(ctx)) return -1;
    const i64 nall = (i64)B.nn + 2 * B.nt;
    PB_PHASE(7, ((size_t)nall + 2) * 4 + 1024);
    B.ed_cnt = PB_ALLOC(7, u32, (size_t)nall + 1);
    PB_RUN(st_edge_count, nall);
    PB_SCAN32(B.ed_cnt, nall);
    u32 tot;
    PB_FETCH(&tot, B.ed_cnt + nall, 4);
    // the counts live in phase 7's buffer; the records go to phase 6 (shared with the pack views)
    PB_PHASE(6, ((size_t)tot + 1) * sizeof(EdgeRec) + 1024);
    B.edges = PB_ALLOC(6, EdgeRec, (size_t)tot + 1);
    PB_RUN(st_edge_fill, nall);
    B.nedges = (i32)tot;
#ifndef PB_HOSTSIM
    CK(ctx_sync(ctx));
#endif
    return 0;
}

int pb200_get_edges(pb200_ctx* ctx, pb200_edge* out) {
    if (!ctx || !ctx->have) return -2;
    if (ctx->B.nedges > 0) PB_TO_HOST(out, ctx->B.edges, (size_t)ctx->B.nedges * sizeof(EdgeRec));
    return 0;
}

int pb200_bellman_ford(pb200_ctx* ctx, int32_t n_nodes, int32_t n_edges, const int32_t* src, const int32_t* dst,
                       const uint32_t* weight_limbs, int32_t source, int32_t target, int32_t* path_out,
                       int32_t* path_len) {
    if (!ctx || n_nodes < 1 || n_edges < 0) return -2;
    DevBuf& b = ctx->scratch;
    size_t need = (size_t)n_edges * (8 + sizeof(WInt)) + (size_t)n_nodes * (sizeof(WInt) + 8) + 8192;
    if (buf_ensure(ctx, b, need)) return -1;
    BFArgs a;
    a.n_nodes = n_nodes;
    a.n_edges = n_edges;
    a.source = source;
    a.target = target;
    i32* dsrc = (i32*)buf_take(b, (size_t)n_edges * 4 + 4);
    i32* ddst = (i32*)buf_take(b, (size_t)n_edges * 4 + 4);
    WInt* dw = (WInt*)buf_take(b, (size_t)n_edges * sizeof(WInt) + 32);
    a.dist = (WInt*)buf_take(b, (size_t)n_nodes * sizeof(WInt));
    a.parent = (i32*)buf_take(b, (size_t)n_nodes * 4);
    a.path = (i32*)buf_take(b, (size_t)n_nodes * 4);
    a.path_len = (i32*)buf_take(b, 8);
    a.src = dsrc;
    a.dst = ddst;
    a.w = dw;
#ifndef PB_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    if (n_edges > 0) {
        CK(cudaMemcpyAsync(dsrc, src, (size_t)n_edges * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ddst, dst, (size_t)n_edges * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(dw, weight_limbs, (size_t)n_edges * sizeof(WInt), cudaMemcpyHostToDevice, ctx->stream));
    }
    k_bf_literal<<<1, 1, 0, ctx->stream>>>(a);
    CK(cudaGetLastError());
#else
    if (n_edges > 0) {
        memcpy(dsrc, src, (size_t)n_edges * 4);
        memcpy(ddst, dst, (size_t)n_edges * 4);
        memcpy(dw, weight_limbs, (size_t)n_edges * sizeof(WInt));
    }
    bf_literal(a);
#endif
    PB_TO_HOST(path_len, a.path_len, 4);
    if (*path_len > 0) PB_TO_HOST(path_out, a.path, (size_t)(*path_len) * 4);
    return 0;
}

// src/phanotate_connect.c:78-121 (get_connected) over the edges (left[i], right[i]) in add_edge order (:62-76)
int pb200_connect(pb200_ctx* ctx, const int32_t* left, const int32_t* right, int32_t n, int32_t* out, int64_t cap_rows,
                  int64_t* n_rows) {
    if (!ctx || n < 0 || !n_rows || (n > 0 && (!left || !right))) return -2;
    *n_rows = 0;
    if (n == 0) return 0;
    if (n > (1 << 22)) {
        ctx->err = "pb200_connect: more than 4,194,304 edges";
        return -2;
    }
    ConnArgs a;
    a.n = n;
    a.nchunk = (n + CN_CHUNK - 1) / CN_CHUNK;
    const i64 items = (i64)n * a.nchunk;
    if (buf_ensure(ctx, ctx->conn, 2 * ((size_t)n * 4 + 256) + (size_t)(items + 1) * 8 + 1024)) return -1;
    i32* dl = (i32*)buf_take(ctx->conn, (size_t)n * 4);
    i32* dr = (i32*)buf_take(ctx->conn, (size_t)n * 4);
    a.cnt = (u64*)buf_take(ctx->conn, (size_t)(items + 1) * 8);
    a.left = dl;
    a.right = dr;
    a.out = nullptr;
    u64 total = 0;
#ifndef PB_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dl, left, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dr, right, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    i64 gx = ((i64)n + CN_BLOCK - 1) / CN_BLOCK;
    if (gx > (i64)ctx->sm_count * 8) gx = (i64)ctx->sm_count * 8;
    const dim3 grid((unsigned)gx, (unsigned)a.nchunk);
    k_connect<false><<<grid, CN_BLOCK, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
#else
    memcpy(dl, left, (size_t)n * 4);
    memcpy(dr, right, (size_t)n * 4);
    for (i64 it = 0; it < items; it++) conn_item(a, it, false);
#endif
    if (dev_scan<u64>(ctx, a.cnt, items)) return -1;
    PB_FETCH(&total, a.cnt + items, 8);
    *n_rows = (int64_t)total;
    if (!out || cap_rows < (int64_t)total || total == 0) return 0;
    if (buf_ensure(ctx, ctx->conn_out, (size_t)total * 8 + 256)) return -1;
    a.out = (i32*)buf_take(ctx->conn_out, (size_t)total * 8);
#ifndef PB_HOSTSIM
    k_connect<true><<<grid, CN_BLOCK, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
#else
    for (i64 it = 0; it < items; it++) conn_item(a, it, true);
#endif
    PB_FETCH(out, a.out, (size_t)total * 8);
    return 0;
}

// ---- host-side text ingest / output for whole batches (no device work, no context needed)
static inline bool fa_blank(unsigned char ch) { return ch == ' ' || ch == '\t' || ch == '\r'; }
static int host_threads(int64_t work_bytes, int64_t min_per_thread = 4 << 20) {
    int t = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("PB200_HOST_THREADS")) t = atoi(e);
    if (t > 32) t = 32;
    const int64_t by_size = work_bytes / min_per_thread;
    if (t > by_size) t = (int)by_size;
    return t < 1 ? 1 : t;
}
// letters -> 4-bit codes, two per byte, low nibble first (a c g t n r y s w k m b v d h, any case; 15 = anything else, which
// the run flags like the letter itself: KeyError in the reference, functions.py:20-24).  Host threads; no context needed.
// `out` holds (n + 1) / 2 bytes.  What pb200_run(..., PB200_INPUT_PACKED4) and pb200_upload_packed4 take: half the bytes
// over the host link.
int64_t pb200_pack4(const uint8_t* bases, int64_t n, uint8_t* out) {
    if (!bases || !out || n < 0) return -2;
    unsigned char code[256];
    memset(code, 15, sizeof code);
    for (int k = 0; k < 15; k++) {
        code[(unsigned char)PB_NIB2CHR[k]] = (unsigned char)k;
        code[(unsigned char)(PB_NIB2CHR[k] - 32)] = (unsigned char)k;
    }
    const int T = host_threads(n, 8 << 20);
    const int64_t pairs = (n + 1) / 2;
    auto work = [&](int t) {
        const int64_t a = pairs / T * t, b = (t == T - 1) ? pairs : pairs / T * (t + 1);
        for (int64_t k = a; k < b; k++) {
            const unsigned lo = code[bases[2 * k]], hi = (2 * k + 1 < n) ? code[bases[2 * k + 1]] : 0u;
            out[k] = (uint8_t)(lo | (hi << 4));
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return pairs;
}

static int64_t fa_count_range(const char* p, const char* end) {
    int64_t rec = 0;
    while (p < end) {
        if (*p == '>') rec++;
        const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!q) break;
        p = q + 1;
    }
    return rec;
}
int64_t pb200_fasta_count(const char* data, int64_t n) {
    const int T = host_threads(n);
    std::vector<const char*> cut(T + 1);
    cut[0] = data;
    cut[T] = data + n;
    for (int t = 1; t < T; t++) {
        const char* p = data + n / T * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char* q = (const char*)memchr(p, '\n', (size_t)(data + n - p));
        cut[t] = q ? q + 1 : data + n;
    }
    std::vector<int64_t> rec(T, 0);
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back([&, t] { rec[t] = fa_count_range(cut[t], cut[t + 1]); });
    rec[0] = fa_count_range(cut[0], cut[1]);
    for (auto& x : th) x.join();
    int64_t r = 0;
    for (int t = 0; t < T; t++) r += rec[t];
    return r;
}
// one pass over the lines of [p, end): counts (bases == nullptr) or writes.  `rec` / `nb` = records and bases before p.
static void fa_lines(const char* data, const char* p, const char* end, bool in_record, uint8_t* bases, int64_t* offsets,
                     int64_t* name_begin, int64_t* name_end, int64_t& rec, int64_t& nb) {
    while (p < end) {
        const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = q ? q : end;                    // line = [p, le)
        if (*p == '>') {
            if (bases) {
                offsets[rec] = nb;
                const char* a = p + 1;
                while (a < le && fa_blank((unsigned char)*a)) a++;
                const char* b = a;
                while (b < le && !fa_blank((unsigned char)*b)) b++;
                name_begin[rec] = a - data;
                name_end[rec] = b - data;
            }
            rec++;
            in_record = true;
        } else if (in_record) {                          // text before the first header is ignored
            const char* e2 = le;
            while (e2 > p && fa_blank((unsigned char)e2[-1])) e2--;          // CRLF / trailing blanks
            const char* a = p;
            while (a < e2 && fa_blank((unsigned char)*a)) a++;
            const size_t len = (size_t)(e2 - a);
            if (len) {
                if (!memchr(a, ' ', len) && !memchr(a, '\t', len) && !memchr(a, '\r', len)) {
                    if (bases) memcpy(bases + nb, a, len);
                    nb += (int64_t)len;
                } else {
                    for (const char* c = a; c < e2; c++)
                        if (!fa_blank((unsigned char)*c)) {
                            if (bases) bases[nb] = (uint8_t)*c;
                            nb++;
                        }
                }
            }
        }
        if (!q) break;
        p = q + 1;
    }
}
// The file is cut into one range of whole lines per host thread; a counting pass gives every range its first record
// and first base, a writing pass fills the packed batch in place (two passes over memory-resident text).
int64_t pb200_fasta_parse(const char* data, int64_t n, uint8_t* bases, int64_t* offsets, int64_t* name_begin,
                          int64_t* name_end, int64_t max_records) {
    const int T = host_threads(n);
    std::vector<const char*> cut(T + 1);
    cut[0] = data;
    cut[T] = data + n;
    for (int t = 1; t < T; t++) {
        const char* p = data + n / T * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char* q = (const char*)memchr(p, '\n', (size_t)(data + n - p));
        cut[t] = q ? q + 1 : data + n;
    }
    std::vector<int64_t> rec(T + 1, 0), nb(T + 1, 0);
    // a range continues a record iff any range before it saw a header: decided after the counting pass, so a range
    // counts its header-less leading lines separately
    std::vector<int64_t> lead(T, 0);
    auto count = [&](int t) {
        // leading lines before the range's first header belong to the previous record (if there is one)
        const char* p = cut[t];
        const char* end = cut[t + 1];
        const char* h = p;
        while (h < end && *h != '>') {
            const char* q = (const char*)memchr(h, '\n', (size_t)(end - h));
            if (!q) {
                h = end;
                break;
            }
            h = q + 1;
        }
        int64_t r0 = 0, b0 = 0;
        fa_lines(data, p, h, true, nullptr, nullptr, nullptr, nullptr, r0, b0);
        lead[t] = b0;
        int64_t r1 = 0, b1 = 0;
        fa_lines(data, h, end, false, nullptr, nullptr, nullptr, nullptr, r1, b1);
        rec[t + 1] = r1;
        nb[t + 1] = b1;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(count, t);
        count(0);
        for (auto& x : th) x.join();
    }
    // prefix: leading lines count only when a record is open
    int64_t r = 0, b = 0;
    std::vector<int64_t> rec0(T), nb0(T);
    std::vector<char> open(T);
    for (int t = 0; t < T; t++) {
        rec0[t] = r;
        nb0[t] = b;
        open[t] = r > 0;
        if (r > 0) b += lead[t];
        r += rec[t + 1];
        b += nb[t + 1];
    }
    if (r > max_records) return -1;
    auto fill = [&](int t) {
        int64_t rr = rec0[t], bb = nb0[t];
        fa_lines(data, cut[t], cut[t + 1], open[t] != 0, bases, offsets, name_begin, name_end, rr, bb);
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(fill, t);
        fill(0);
        for (auto& x : th) x.join();
    }
    offsets[r] = b;
    return r;
}
// "%E" of a double (what Locus.tabular prints: '%E' % weight, phanotate.py:75-76), exactly as printf rounds it -- half-even
// on the exact binary value -- in 128-bit integer arithmetic for 1e-16 <= |x| < 1.7e38 (every score but the astronomic
// ones), sprintf otherwise.  ~15x faster than glibc's printf_fp, which was most of the tabular writer's time.
static int fmt_E(char* w, double x) {
    typedef unsigned __int128 u128;
    static const u128 P10[39] = {
        (u128)1ull, (u128)10ull, (u128)100ull, (u128)1000ull, (u128)10000ull, (u128)100000ull, (u128)1000000ull, (u128)10000000ull,
        (u128)100000000ull, (u128)1000000000ull, (u128)10000000000ull, (u128)100000000000ull, (u128)1000000000000ull,
        (u128)10000000000000ull, (u128)100000000000000ull, (u128)1000000000000000ull, (u128)10000000000000000ull,
        (u128)100000000000000000ull, (u128)1000000000000000000ull, (u128)10000000000000000000ull,
        (u128)10000000000000000000ull * 10u, (u128)10000000000000000000ull * 100u, (u128)10000000000000000000ull * 1000u,
        (u128)10000000000000000000ull * 10000u, (u128)10000000000000000000ull * 100000u, (u128)10000000000000000000ull * 1000000u,
        (u128)10000000000000000000ull * 10000000u, (u128)10000000000000000000ull * 100000000u,
        (u128)10000000000000000000ull * 1000000000u, (u128)10000000000000000000ull * 10000000000ull,
        (u128)10000000000000000000ull * 100000000000ull, (u128)10000000000000000000ull * 1000000000000ull,
        (u128)10000000000000000000ull * 10000000000000ull, (u128)10000000000000000000ull * 100000000000000ull,
        (u128)10000000000000000000ull * 1000000000000000ull, (u128)10000000000000000000ull * 10000000000000000ull,
        (u128)10000000000000000000ull * 100000000000000000ull, (u128)10000000000000000000ull * 1000000000000000000ull,
        (u128)10000000000000000000ull * 10000000000000000000ull};
    const double ax = x < 0 ? -x : x;
    if (!(ax >= 1e-16 && ax < 1.7e38)) return sprintf(w, "%E", x);
    int e2;
    const double fr = frexp(ax, &e2);                       // ax = fr * 2^e2, 0.5 <= fr < 1
    const uint64_t m = (uint64_t)ldexp(fr, 53);             // exact: 53-bit integer
    const int e = e2 - 53;                                  // ax = m * 2^e
    int E = (int)floor(log10(ax));
    uint64_t q = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        const int k = E - 6;                                // q = round(ax / 10^k)
        u128 num, den;
        if (k >= 0) {
            if (e >= 0) {
                num = (u128)m << e;
                den = P10[k];
            } else {
                num = (u128)m;
                den = P10[k] << (-e);
            }
        } else {
            if (-k > 22 || e >= 0) return sprintf(w, "%E", x);
            num = (u128)m * P10[-k];
            den = (u128)1 << (-e);
        }
        const u128 qq = num / den, rem = num - qq * den;
        q = (uint64_t)qq;
        if (qq < 1000000u) {                                // the estimate of E was one too high
            E--;
            continue;
        }
        if (qq >= 10000000u) {
            E++;
            continue;
        }
        const u128 twice = rem << 1;
        if (twice > den || (twice == den && (q & 1))) q++;
        if (q == 10000000u) {
            q = 1000000u;
            E++;
        }
        break;
    }
    char* p = w;
    if (x < 0) *p++ = '-';
    char dg[8];
    for (int i = 6; i >= 0; i--) {
        dg[i] = (char)('0' + q % 10);
        q /= 10;
    }
    *p++ = dg[0];
    *p++ = '.';
    for (int i = 1; i < 7; i++) *p++ = dg[i];
    *p++ = 'E';
    int ae = E < 0 ? -E : E;
    *p++ = E < 0 ? '-' : '+';
    if (ae >= 100) *p++ = (char)('0' + ae / 100);
    *p++ = (char)('0' + (ae / 10) % 10);
    *p++ = (char)('0' + ae % 10);
    *p = 0;
    return (int)(p - w);
}
// exported for the tests: the text fmt_E writes for x (at most 31 bytes + NUL)
int pb200_format_score(double x, char* out) { return fmt_E(out, x); }
static inline char* fmt_int(char* w, int v) {
    char t[12];
    int n = 0;
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    do {
        t[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *w++ = '-';
    while (n) *w++ = t[--n];
    return w;
}
// Locus.tabular (locus.py:39-56) for every contig of a batch: "#id:" / "#START..." header, then one row per call with
// left/right swapped on the reverse strand and the score printed with %E.  Returns the bytes written or -(bytes needed).
// Contig ranges are formatted by the host threads into the (upper-bound sized) output and then closed up.
int64_t pb200_format_tabular(const pb200_call* calls, const pb200_contig* contigs, int32_t n_contigs, const char* names,
                             const int64_t* name_off, char* out, int64_t cap) {
    std::vector<int64_t> at((size_t)n_contigs + 1, 0);
    for (int32_t k = 0; k < n_contigs; k++)
        at[k + 1] = at[k] + 48 + (name_off[k + 1] - name_off[k]) +
                    (int64_t)contigs[k].n_calls * (64 + (name_off[k + 1] - name_off[k]));
    const int64_t need = at[n_contigs];
    if (need > cap) return -need;
    const int T = host_threads(need, 64 << 10);          // (formatting costs ~30 ns per byte: small pieces pay)
    std::vector<int32_t> cut(T + 1, n_contigs);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        const int64_t want = need / T * t;
        cut[t] = (int32_t)(std::lower_bound(at.begin(), at.end(), want) - at.begin());
        if (cut[t] > n_contigs) cut[t] = n_contigs;
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    std::vector<int64_t> len(T, 0);
    auto fmt = [&](int t) {
        char* w = out + at[cut[t]];
        char* const w0 = w;
        for (int32_t k = cut[t]; k < cut[t + 1]; k++) {
            const char* nm = names + name_off[k];
            const int nl = (int)(name_off[k + 1] - name_off[k]);
            w += sprintf(w, "#id:\t%.*s\n#START\tSTOP\tFRAME\tCONTIG\tSCORE\n", nl, nm);
            const pb200_call* c = calls + contigs[k].call_off;
            for (int32_t i = 0; i < contigs[k].n_calls; i++) {
                if (c[i].strand == 2 || c[i].strand == -2) continue;       // a tRNA on the path: Locus.tabular lists CDS only (locus.py:41)
                const int fwd = c[i].strand > 0;
                w = fmt_int(w, fwd ? c[i].left : c[i].right);
                *w++ = '\t';
                w = fmt_int(w, fwd ? c[i].right : c[i].left);
                *w++ = '\t';
                *w++ = fwd ? '+' : '-';
                *w++ = '\t';
                memcpy(w, nm, (size_t)nl);
                w += nl;
                *w++ = '\t';
                w += fmt_E(w, c[i].score);
                *w++ = '\n';
            }
        }
        len[t] = w - w0;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(fmt, t);
        fmt(0);
        for (auto& x : th) x.join();
    }
    int64_t total = len[0];
    for (int t = 1; t < T; t++) {                        // close the gaps between the threads' blocks
        memmove(out + total, out + at[cut[t]], (size_t)len[t]);
        total += len[t];
    }
    return total;
}

int pb200_stage_times(pb200_ctx* ctx, const char** names, float* ms, int cap) {
    if (!ctx) return -2;
#ifndef PB_HOSTSIM
    int n = 0;
    for (auto& t : ctx->times) {
        if (n >= cap) break;
        float v = 0.f;
        if (cudaEventElapsedTime(&v, t.a, t.b) != cudaSuccess) v = -1.f;
        names[n] = t.name;
        ms[n] = v;
        n++;
    }
    return n;
#else
    (void)names;
    (void)ms;
    (void)cap;
    return 0;
#endif
}

// idle / untimed device time in front of every timed stage of the last run (from the end of the previous timed stage,
// or from the start of the run): prefix scans, memsets, host round trips for table sizes, launch latency
int pb200_stage_gaps(pb200_ctx* ctx, const char** names, float* ms, int cap) {
    if (!ctx) return -2;
#ifndef PB_HOSTSIM
    int n = 0;
    cudaEvent_t prev = ctx->run_a;
    for (auto& t : ctx->times) {
        if (n >= cap) break;
        float v = 0.f;
        if (!prev || cudaEventElapsedTime(&v, prev, t.a) != cudaSuccess) v = -1.f;
        names[n] = t.name;
        ms[n] = v;
        prev = t.b;
        n++;
    }
    return n;
#else
    (void)names;
    (void)ms;
    (void)cap;
    return 0;
#endif
}

int pb200_launch_count(pb200_ctx* ctx) { return ctx ? ctx->launches : -2; }

#include "comm.inc"

}  // extern "C"
