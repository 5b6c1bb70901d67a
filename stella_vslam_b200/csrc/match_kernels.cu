// match_kernels.cu -- 256-bit Hamming matchers on sm_100a.
//
// Reference path:
//   match::compute_descriptor_distance_32      src/stella_vslam/match/base.h:20-41
//   match::robust::brute_force_match           src/stella_vslam/match/robust.cc:232-328
//   util::angle::diff                          src/stella_vslam/util/angle.cc:7-16
//
// brute_force_match is sequential in the reference: keyframe keypoints idx_2 are visited in order and every accepted
// match removes its frame keypoint idx_1 from all later searches.  Restated here as
//   (1) a fully parallel pass that keeps, per idx_2, the K smallest (distance, idx_1) keys over the orientation-gated
//       frame keypoints (uint4 loads of the descriptors, __popc on the XOR), and
//   (2) an in-order resolve (one warp per problem) that walks each list skipping taken idx_1; when a list cannot decide
//       the outcome exactly (too many of its entries were taken) the warp recomputes that row against the live set.
// Both steps are exact, so the match pairs equal the reference's bit for bit.
#include <algorithm>
#include <climits>
#include <cmath>
#include <type_traits>
#include <new>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "track_chain.cuh"

namespace b200 {
namespace match {

constexpr int kTopK = 8;            // candidates kept per keyframe keypoint
constexpr int kRowsPerBlock = 64;   // one keyframe keypoint per thread (small CTAs: a 2000-row problem still yields 32 of them)
constexpr int kChunk = 256;         // frame descriptors staged per shared-memory tile (8 KB)
constexpr unsigned kInfKey = 0xFFFFFFFFu;
constexpr unsigned kSentinelIdx = 0x3FFFFFu;  // index no keypoint can have (frames hold < 2^22 - 1 keypoints)
constexpr int kThrLow = 50;         // HAMMING_DIST_THR_LOW  (match/base.h:15)
constexpr int kMaxDist = 256;       // MAX_HAMMING_DIST      (match/base.h:17)

// key = distance << 22 | idx_1 : unsigned order == (distance, then first index) == the reference's strict '<' scan
__device__ __forceinline__ unsigned make_key(unsigned dist, unsigned idx) { return (dist << 22) | idx; }
__device__ __forceinline__ unsigned key_dist(unsigned key) { return key >> 22; }
__device__ __forceinline__ unsigned key_idx(unsigned key) { return key & 0x3FFFFFu; }

// util/angle.cc:7-16 compares and adds in double (float operands promoted).  With float operands these are the same decisions and the
// same values in float arithmetic: -180, 180, 360 and 30 are exact floats, promoting a float is exact (so the comparisons agree), and
// (float)((double)ret + 360.0) rounds an exactly representable double sum once -- which is what __fadd_rn(ret, 360.0f) does.  Float
// keeps the test off the fp64 pipe (it sits in the selection path of every matcher).
__device__ __forceinline__ float angle_diff(float a1, float a2) {
    float ret = __fsub_rn(a1, a2);
    if (ret <= -180.0f) ret = __fadd_rn(ret, 360.0f);
    if (ret > 180.0f) ret = __fsub_rn(ret, 360.0f);
    return ret;
}
__device__ __forceinline__ bool orientation_rejects(float a1, float a2) {  // robust.cc:279
    return fabsf(angle_diff(a1, a2)) > 30.0f;
}

__device__ __forceinline__ unsigned hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x)
           + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---------------------------------------------------------------------------------------------------------------
// all-pairs distance matrix (diagnostics / landmark::compute_descriptor-style consumers)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hamming_matrix_kernel(const uint4* __restrict__ d1, int n1, const uint4* __restrict__ d2, int n2,
                                                             unsigned short* __restrict__ out) {
    __shared__ uint4 s2[64 * 2];
    const int j0 = blockIdx.x * 64, i = blockIdx.y * blockDim.x + threadIdx.x;
    for (int t = threadIdx.x; t < 128; t += blockDim.x) s2[t] = (j0 + t / 2 < n2) ? d2[(size_t)j0 * 2 + t] : make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (i >= n1) return;
    const uint4 a0 = d1[(size_t)i * 2], a1 = d1[(size_t)i * 2 + 1];
    for (int j = 0; j < 64 && j0 + j < n2; ++j) out[(size_t)i * n2 + j0 + j] = (unsigned short)hamming256(a0, a1, s2[2 * j], s2[2 * j + 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// (1) top-K candidate lists.  grid = (ceil(max_n2 / 128), n_problems); thread = one keyframe keypoint idx_2.
// ---------------------------------------------------------------------------------------------------------------
struct Side {                       // one side of the problems: descriptors, strided angles, per-problem (offset, count)
    const uint4* desc;
    const unsigned char* angle;     // angle of keypoint i = *(const float*)(angle + i * angle_stride)
    long long angle_stride;
    const int* off;
    const int* cnt;
};
__device__ __forceinline__ float side_angle(const Side& s, int i) {
    return *reinterpret_cast<const float*>(s.angle + (long long)i * s.angle_stride);
}

__global__ void __launch_bounds__(kRowsPerBlock) topk_kernel(Side S1, Side S2, const unsigned char* __restrict__ valid2,
                                                            int check_orientation, unsigned* __restrict__ lists) {
    __shared__ uint4 s1[kChunk * 2];
    __shared__ float sa[kChunk];
    const uint4* __restrict__ desc1 = S1.desc;
    const uint4* __restrict__ desc2 = S2.desc;
    const int p = blockIdx.y;
    const int b1 = S1.off[p], n1 = S1.cnt[p];
    const int b2 = S2.off[p], n2 = S2.cnt[p];
    if ((int)(blockIdx.x * kRowsPerBlock) >= n2) return;
    const int row = blockIdx.x * kRowsPerBlock + threadIdx.x;
    const bool active = row < n2 && (!valid2 || valid2[b2 + row]);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    float qa = 0.f;
    if (active) {
        q0 = desc2[(size_t)(b2 + row) * 2];
        q1 = desc2[(size_t)(b2 + row) * 2 + 1];
        qa = side_angle(S2, b2 + row);
    }
    unsigned top[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; ++k) top[k] = kInfKey;
    for (int c0 = 0; c0 < n1; c0 += kChunk) {
        const int cn = min(kChunk, n1 - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < cn * 2; t += blockDim.x) s1[t] = desc1[(size_t)(b1 + c0) * 2 + t];
        for (int t = threadIdx.x; t < cn; t += blockDim.x) sa[t] = side_angle(S1, b1 + c0 + t);
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int j = 0; j < cn; ++j) {
            const unsigned dist = hamming256(q0, q1, s1[2 * j], s1[2 * j + 1]);
            const unsigned key = make_key(dist, (unsigned)(c0 + j));
            if (key < top[kTopK - 1]) {
                if (check_orientation && orientation_rejects(sa[j], qa)) continue;
                top[kTopK - 1] = key;
#pragma unroll
                for (int k = kTopK - 1; k > 0; --k) {
                    if (top[k] < top[k - 1]) {
                        const unsigned t = top[k];
                        top[k] = top[k - 1];
                        top[k - 1] = t;
                    }
                }
            }
        }
    }
    if (row < n2) {
        unsigned* o = lists + ((size_t)p * gridDim.x * kRowsPerBlock + row) * kTopK;
#pragma unroll
        for (int k = 0; k < kTopK; ++k) o[k] = top[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// (1b) top-K candidate lists on the 5th-generation tensor cores.  The Hamming distance of two 256-bit descriptors is a dot
//      product in disguise: with the bits mapped to +-1,  popcount(a ^ b) = (256 - <a, b>) / 2  -- exact in int32 -- so the all-pairs
//      distance matrix of a (frame, keyframe) pair is a 2000 x 2000 x 256 int8 GEMM.  One CTA owns 128 keyframe rows (one M tile,
//      expanded once into shared memory in the canonical K-major no-swizzle UMMA layout: 8 x 16-byte core matrices) and walks the
//      frame's keypoints in chunks of 128 (the B operand, expanded by the same threads): ONE thread issues
//      tcgen05.mma.cta_group::1.kind::i8 (M 128, N 128, K 32; 8 per chunk), the accumulators live in TMEM (2 buffers x 128 columns;
//      two CTAs per SM use all 512) and tcgen05.commit signals an mbarrier.  While the tensor core works on chunk c the four warps
//      read chunk c-1 back with tcgen05.ld (thread = keyframe row = TMEM lane) and keep the smallest (distance, index) keys that
//      pass the orientation gate -- the selection of topk_kernel up to a distance cap that cannot change a decision.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTcRows = 128;        // keyframe rows per CTA (one M tile); two CTAs share an SM: one stages / waits while the other selects
constexpr int kTcChunk = 128;       // frame keypoints per B chunk (= N of the MMA)
constexpr int kTcThreads = 256;     // two warps per TMEM lane quarter: each takes one half of a chunk's 128 columns
constexpr unsigned kTcTmemCols = 256;  // two accumulator buffers of 128 columns (two resident CTAs use all 512)
constexpr int kTcTileBytes = 128 * 256;  // one 128-row operand tile of +-1 bytes
struct TcSmem {
    unsigned char a[kTcTileBytes];
    unsigned char b[2][kTcTileBytes];
    uint2 lut[256];                 // descriptor byte -> 8 bytes of +-1
    float ang[3][kTcChunk];         // three: the epilogue of chunk c-1 may still read its angles while chunk c+1 is being staged
    unsigned long long bar[2];
    unsigned tmem_base;
    unsigned pad;
};
__device__ __forceinline__ unsigned long long umma_desc_k_major(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46) (all >> 4), version 1 at
    // [46,48), base offset 0, layout type SWIZZLE_NONE (0) at [61,64).  K-major canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units.
    return (unsigned long long)((smem_addr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16)
           | ((unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_mbar_wait(unsigned long long* bar, unsigned parity) {
    // bounded: a tensor-core fault must surface as a launch error, never as a hung GPU
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    for (unsigned spin = 0; spin < (1u << 28); ++spin) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
// +-1 expansion of one 16-byte half descriptor (8 K-chunks of 16 bytes) into the operand tile: row r of a 128-row tile, chunks c0..c0+7
__device__ __forceinline__ void tc_expand_half(unsigned char* tile, const uint2* lut, int r, int c0, uint4 bits) {
    unsigned char* dst = tile + (r >> 3) * 128 + (r & 7) * 16;
    const unsigned w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // chunk c0 + i <- descriptor bytes 2i, 2i+1 of this half
        const unsigned hw = (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        const uint2 lo = lut[hw & 0xFFu], hi = lut[hw >> 8];
        *reinterpret_cast<uint4*>(dst + (c0 + i) * 2048) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
}

__global__ void __launch_bounds__(kTcThreads, 2) topk_tc_kernel(Side S1, Side S2, const unsigned char* __restrict__ valid2, int check_orientation,
                                                                unsigned* __restrict__ lists, int list_rows, unsigned cap) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(tc_smem_raw);
    const uint4* __restrict__ desc1 = S1.desc;
    const uint4* __restrict__ desc2 = S2.desc;
    const int p = blockIdx.y;
    const int b1 = S1.off[p], n1 = S1.cnt[p];
    const int b2 = S2.off[p], n2 = S2.cnt[p];
    const int row0 = blockIdx.x * kTcRows;
    if (row0 >= n2) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- one-time setup: TMEM, barriers, the byte -> +-1 table, the A operand (this CTA's keyframe rows)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&sm.tmem_base)), "r"(kTcTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&sm.bar[0])), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&sm.bar[1])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        for (int e = tid; e < 256; e += kTcThreads) {  // bit i of the byte -> byte i: +1 (0x01) if set, -1 (0xFF) if clear
            unsigned lo = 0, hi = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                lo |= (((e >> i) & 1) ? 0x01u : 0xFFu) << (8 * i);
                hi |= (((e >> (4 + i)) & 1) ? 0x01u : 0xFFu) << (8 * i);
            }
            sm.lut[e] = make_uint2(lo, hi);
        }
    }
    __syncthreads();
    // thread = (keyframe row = TMEM lane, column half): warps 0-3 select from columns 0-63 of every chunk, warps 4-7 from 64-127 (a warp
    // may only read the TMEM lanes of its quarter, warp % 4); the two half lists of a row are merged at the end.  Eight resident warps
    // per CTA instead of four: the selection pass is latency-bound (LDTM, dependent compare chains), not issue-bound.
    const int rtid = tid & 127, half = tid >> 7;
    const int my_row = row0 + rtid;
    const bool active = my_row < n2 && (!valid2 || valid2[b2 + my_row]);
    {
        // A operand: each thread expands ONE 16-byte half of its row's descriptor (8 of the 16 K-chunks)
        const bool real = my_row < n2;
        unsigned char* tile = sm.a;
        if (real) {
            tc_expand_half(tile, sm.lut, rtid, 8 * half, desc2[(size_t)(b2 + my_row) * 2 + half]);
        } else {  // rows past the keyframe: zeros (their lists are never written)
            unsigned char* dst = tile + (rtid >> 3) * 128 + (rtid & 7) * 16;
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + (8 * half + c) * 2048) = make_uint4(0, 0, 0, 0);
        }
    }
    const float qa = active ? side_angle(S2, b2 + my_row) : 0.f;
    unsigned top[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; ++k) top[k] = kInfKey;
    // Distance cap.  robust.cc:297-308 accepts a match only if best <= 50 and lowe * second >= best, so a second-best distance matters
    // only while lowe * second < 50: a candidate farther than cap = ceil(50 / lowe) can change no decision and is not listed (random
    // descriptor pairs are ~128 +- 8 apart, so this removes nearly every insertion).  Rows without a landmark list nothing.
    unsigned thr = active ? make_key(cap + 1u, 0u) : 0u;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 at [4,6)), A = B = signed int8 (1 at [7,10) and [10,13)), both K-major,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kTcChunk >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned a_addr = (unsigned)__cvta_generic_to_shared(sm.a);
    const unsigned b_addr[2] = {(unsigned)__cvta_generic_to_shared(sm.b[0]), (unsigned)__cvta_generic_to_shared(sm.b[1])};
    const int n_chunks = (n1 + kTcChunk - 1) / kTcChunk;
    unsigned tmem = 0;
    // the descriptor bits (and angle) of this thread's row of the NEXT chunk are fetched one iteration ahead: the L2 latency hides behind
    // the selection pass instead of standing in front of the tensor core
    uint4 nb = make_uint4(0, 0, 0, 0);
    float nang = 0.f;
    if (rtid < n1) {
        nb = desc1[(size_t)(b1 + rtid) * 2 + half];
        if (half == 0) nang = side_angle(S1, b1 + rtid);
    }
    for (int c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {
            // B operand of chunk c: 128 frame keypoints, thread = row
            const int buf = c & 1, r = rtid, j = c * kTcChunk + r;
            if (j < n1) {
                tc_expand_half(sm.b[buf], sm.lut, r, 8 * half, nb);
                if (half == 0) sm.ang[c % 3][r] = nang;
            } else {
                unsigned char* dst = sm.b[buf] + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) *reinterpret_cast<uint4*>(dst + (8 * half + cc) * 2048) = make_uint4(0, 0, 0, 0);
            }
            const int jn = j + kTcChunk;
            if (jn < n1) {
                nb = desc1[(size_t)(b1 + jn) * 2 + half];
                if (half == 0) nang = side_angle(S1, b1 + jn);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's async proxy
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = sm.tmem_base;
        if (c < n_chunks && tid == 0) {
            const int buf = c & 1;
            const unsigned d_col = tmem + (unsigned)(buf * kTcChunk);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)  // K = 32 bytes per instruction = chunks 2 ks, 2 ks + 1 (2048 bytes apart)
                tc_mma_i8(d_col, umma_desc_k_major(a_addr + ks * 4096, 2048, 128), umma_desc_k_major(b_addr[buf] + ks * 4096, 2048, 128), idesc,
                          ks > 0 ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&sm.bar[buf]))
                         : "memory");
        }
        if (c > 0) {
            // epilogue of chunk c - 1 while the tensor core runs chunk c
            const int e = c - 1, buf = e & 1, c0 = e * kTcChunk;
            tc_mbar_wait(&sm.bar[buf], (unsigned)(e >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(buf * kTcChunk + half * (kTcChunk / 2));
            // key = distance << 22 | index with distance = (256 - dot) / 2:  (256 - dot) << 21 has bit 21 clear (the dot product of two
            // +-1 vectors of even length is even), so the key is one multiply-add.  Four keys are tested against the row's threshold with
            // one 3-input minimum, one compare and one warp vote; only a group that holds a candidate for some row of the warp inserts.
            // The threshold starts at the distance cap (see below), so candidates are rare: a warp's 32 rows see a few dozen in all.
            const int ch0 = c0 + half * (kTcChunk / 2);  // first frame keypoint of this thread's half of the chunk
            const unsigned kbase = (256u << 21) + (unsigned)ch0;
            auto scan = [&](auto full_chunk) {
#pragma unroll 1
                for (int g = 0; g < kTcChunk / 64; ++g) {
                    if (!decltype(full_chunk)::value && ch0 + 32 * g >= n1) break;
                    unsigned v[32];
                    tc_ld32(taddr + 32 * g, v);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        unsigned key[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            key[u] = (unsigned)((int)v[i + u] * -(1 << 21)) + (kbase + (unsigned)(32 * g + i + u));
                            if (!decltype(full_chunk)::value && ch0 + 32 * g + i + u >= n1) key[u] = kInfKey;  // zero padding of the last chunk
                        }
                        const unsigned m4 = min(__vimin3_u32(key[0], key[1], key[2]), key[3]);
                        if (__any_sync(0xFFFFFFFFu, m4 < thr)) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (key[u] < thr) {
                                    if (check_orientation && orientation_rejects(sm.ang[e % 3][half * (kTcChunk / 2) + 32 * g + i + u], qa)) continue;
                                    top[kTopK - 1] = key[u];
#pragma unroll
                                    for (int k = kTopK - 1; k > 0; --k) {
                                        if (top[k] < top[k - 1]) {
                                            const unsigned t2 = top[k];
                                            top[k] = top[k - 1];
                                            top[k - 1] = t2;
                                        }
                                    }
                                    thr = min(thr, top[kTopK - 1]);
                                }
                            }
                        }
                    }
                }
            };
            if (c0 + kTcChunk <= n1) scan(std::true_type{});
            else scan(std::false_type{});
        }
    }
    // merge the two half lists of every row: the upper half parks its (sorted) keys in the B staging area, the lower half inserts them
    __syncthreads();  // (nobody reads the operand tiles any more: the last MMA was committed and waited for)
    unsigned* park = reinterpret_cast<unsigned*>(sm.b[0]);
    if (half == 1) {
#pragma unroll
        for (int k = 0; k < kTopK; ++k) park[k * 128 + rtid] = top[k];
    }
    __syncthreads();
    if (half == 0) {
#pragma unroll
        for (int k2 = 0; k2 < kTopK; ++k2) {
            const unsigned key = park[k2 * 128 + rtid];
            if (key >= top[kTopK - 1]) continue;  // (both lists ascending: nothing after this one can enter either, but the unrolled form is branch-light)
            top[kTopK - 1] = key;
#pragma unroll
            for (int k = kTopK - 1; k > 0; --k) {
                if (top[k] < top[k - 1]) {
                    const unsigned t2 = top[k];
                    top[k] = top[k - 1];
                    top[k - 1] = t2;
                }
            }
        }
        // lists shorter than 8: the slots name no candidate but state the bound "everything else is farther than the cap"
#pragma unroll
        for (int k = 0; k < kTopK; ++k)
            if (top[k] == kInfKey && cap < (unsigned)kMaxDist) top[k] = make_key(cap + 1u, kSentinelIdx);
        if (my_row < n2) {
            unsigned* o = lists + ((size_t)p * list_rows + my_row) * kTopK;
#pragma unroll
            for (int k = 0; k < kTopK; ++k) o[k] = top[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// (2) in-order resolve + compaction.  One warp per problem.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned warp_min(unsigned v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, s));
    return v;
}

// exact best/second over the live (not taken, orientation-gated) frame keypoints: the reference's inner loop, one warp.
// desc1/angle1 point either to the problem's frame side in global memory (angle stride in bytes) or to its staged copy in
// shared memory.
__device__ void exact_row(const uint4* desc1, const unsigned char* angle1, long long angle_stride, int n1, const volatile unsigned* taken,
                          uint4 q0, uint4 q1, float qa, int check_orientation, int lane, unsigned* best_key, unsigned* second_dist) {
    unsigned k1 = kInfKey, k2 = kInfKey;  // two smallest keys seen by this lane
    for (int i0 = lane; i0 < n1; i0 += 128) {
        uint4 a0[4], a1[4];
        float ang[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // four independent loads in flight per lane
            const int i = min(i0 + 32 * u, n1 - 1);
            a0[u] = desc1[(size_t)i * 2];
            a1[u] = desc1[(size_t)i * 2 + 1];
            ang[u] = *reinterpret_cast<const float*>(angle1 + (long long)i * angle_stride);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            if (i >= n1) continue;
            if ((taken[i >> 5] >> (i & 31)) & 1u) continue;
            if (check_orientation && orientation_rejects(ang[u], qa)) continue;
            const unsigned key = make_key(hamming256(q0, q1, a0[u], a1[u]), (unsigned)i);
            if (key < k1) {
                k2 = k1;
                k1 = key;
            } else if (key < k2) {
                k2 = key;
            }
        }
    }
    const unsigned b = warp_min(k1);
    const unsigned mine = (k1 == b) ? k2 : k1;  // keys are unique (they embed idx_1), so exactly one lane owns b
    const unsigned s = warp_min(mine);
    *best_key = b;
    *second_dist = (s == kInfKey) ? (unsigned)kMaxDist : key_dist(s);
}

// Decision of one keyframe keypoint from its sorted candidate list against the current `taken` set.
//   returns 0: decided, no match; 1: decided, match with key *best; 2: undecidable from the list (exact row scan needed)
__device__ __forceinline__ int decide_row(const unsigned (&keys)[kTopK], bool full, unsigned tail_dist, const volatile unsigned* taken, float lowe_ratio,
                                          unsigned* best) {
    unsigned best_key = kInfKey, second_dist = kMaxDist;
    int n_live = 0;
#pragma unroll
    for (int k = 0; k < kTopK; ++k) {
        const unsigned key = keys[k];
        if (key == kInfKey) continue;
        const unsigned i1 = key_idx(key);
        if ((taken[i1 >> 5] >> (i1 & 31)) & 1u) continue;
        if (n_live == 0) best_key = key;
        else if (n_live == 1) second_dist = key_dist(key);
        ++n_live;
    }
    // full: unlisted candidates exist; every unlisted distance is >= tail_dist
    if (full && n_live < 2) {
        if (n_live == 1) {
            // best is exact; the true second distance lies in [tail_dist, 256]
            const unsigned bd = key_dist(best_key);
            if (bd > (unsigned)kThrLow) return 0;
            if (__fmul_rn(lowe_ratio, (float)tail_dist) < (float)bd) return 2;
            second_dist = tail_dist;  // passes the ratio test even with the smallest possible second distance
        } else {
            // every listed candidate is taken: the best live distance is >= tail_dist
            return (tail_dist > (unsigned)kThrLow) ? 0 : 2;
        }
    }
    if (best_key == kInfKey) return 0;
    const unsigned bd = key_dist(best_key);
    // robust.cc:297-308: threshold, then Lowe ratio in float
    if (bd > (unsigned)kThrLow || __fmul_rn(lowe_ratio, (float)second_dist) < (float)bd) return 0;
    *best = best_key;
    return 1;
}

// One warp per problem.  Rows (keyframe keypoints) are visited in batches of 32, one per lane.  Every lane decides its row
// against the taken set as of the last commit; a lane's decision is final iff no earlier, not yet committed lane accepts a
// frame keypoint that appears in its candidate list -- so the longest conflict-free prefix of the batch is committed at
// once and the rest re-decided.  The result is identical to the reference's strictly sequential loop.
__global__ void __launch_bounds__(32) resolve_kernel(Side S1, Side S2, const unsigned char* __restrict__ valid2,
                                                     const unsigned* __restrict__ lists, float lowe_ratio, int check_orientation,
                                                     int* __restrict__ matched, unsigned* __restrict__ taken_g, int taken_words,
                                                     int list_rows, int matched_stride, int pairs_stride,
                                                     int* __restrict__ pairs_out, int* __restrict__ n_pairs, int use_smem) {
    const uint4* __restrict__ desc1 = S1.desc;
    const uint4* __restrict__ desc2 = S2.desc;
    extern __shared__ __align__(16) unsigned resolve_smem[];  // [desc1 copy 32 B x n][angles][taken bitmap][idx_1 -> idx_2 table][claim table], as far as they fit
    const int p = blockIdx.x, lane = threadIdx.x;
    const int b1 = S1.off[p], n1 = S1.cnt[p];
    const int b2 = S2.off[p], n2 = S2.cnt[p];
    // the sequential state lives in shared memory (30-cycle instead of L2 latency on the critical path); very large frames fall
    // back to the global scratch
    // use_smem: 0 = everything in global scratch, 1 = taken + table in shared memory, 2 = additionally the frame descriptors and
    // angles (the exact fallback then scans shared memory instead of L2)
    const int stage_words = (use_smem == 2) ? 9 * matched_stride : 0;  // 8 words of descriptor + 1 angle per keypoint
    unsigned* taken = use_smem ? resolve_smem + stage_words : taken_g + (size_t)p * taken_words;
    int* m21 = use_smem ? reinterpret_cast<int*>(resolve_smem + stage_words + taken_words) : matched + (size_t)p * matched_stride;
    int* claim = use_smem ? reinterpret_cast<int*>(resolve_smem + stage_words + taken_words + matched_stride) : nullptr;
    int* pairs = pairs_out + 2 * (size_t)p * pairs_stride;
    lists += (size_t)p * list_rows * kTopK;
    for (int i = lane; i < (n1 + 31) / 32; i += 32) taken[i] = 0u;
    for (int i = lane; i < n1; i += 32) m21[i] = -1;
    if (claim)
        for (int i = lane; i < n1; i += 32) claim[i] = 255;
    __syncwarp();
    const uint4* d1 = desc1 + (size_t)b1 * 2;
    const unsigned char* a1p = S1.angle + (long long)b1 * S1.angle_stride;
    long long a1s = S1.angle_stride;
    if (use_smem == 2) {
        uint4* sd = reinterpret_cast<uint4*>(resolve_smem);
        float* sa = reinterpret_cast<float*>(resolve_smem + 8 * matched_stride);
        for (int i = lane; i < 2 * n1; i += 32) sd[i] = d1[i];
        for (int i = lane; i < n1; i += 32) sa[i] = side_angle(S1, b1 + i);
        d1 = sd;
        a1p = reinterpret_cast<const unsigned char*>(sa);
        a1s = sizeof(float);
        __syncwarp();
    }
    for (int base = 0; base < n2; base += 32) {
        const int r = base + lane;
        const bool has_row = r < n2 && (!valid2 || valid2[b2 + r]);  // robust.cc:255-262
        unsigned keys[kTopK];
        {
            const uint4* lp = reinterpret_cast<const uint4*>(lists + (size_t)min(r, n2 - 1) * kTopK);
            const uint4 k0 = lp[0], k1 = lp[1];
            keys[0] = k0.x; keys[1] = k0.y; keys[2] = k0.z; keys[3] = k0.w;
            keys[4] = k1.x; keys[5] = k1.y; keys[6] = k1.z; keys[7] = k1.w;
        }
        // A list is "full" when candidates exist that it does not name: its last entry then bounds their distance from below.  The
        // tensor-core lister only names candidates up to a distance cap and pads with sentinel keys (cap + 1, impossible index).
        const bool list_full = keys[kTopK - 1] != kInfKey;
        const unsigned tail_dist = key_dist(keys[kTopK - 1]);
#pragma unroll
        for (int k = 0; k < kTopK; ++k)
            if (key_idx(keys[k]) == kSentinelIdx) keys[k] = kInfKey;
        unsigned pending = __ballot_sync(0xFFFFFFFFu, has_row);
        while (pending) {
            const bool mine = (pending >> lane) & 1u;
            unsigned best = kInfKey;
            const int st = mine ? decide_row(keys, list_full, tail_dist, taken, lowe_ratio, &best) : 0;
            const unsigned accept_mask = __ballot_sync(0xFFFFFFFFu, mine && st == 1);
            const unsigned exact_mask = __ballot_sync(0xFFFFFFFFu, mine && st == 2);
            // conflicts with earlier accepting lanes of this round
            bool conflict = false;
            const unsigned my_best_idx = key_idx(best);
            if (claim) {
                // claim[i] = lowest lane that accepts frame keypoint i this round; a lane conflicts when a lower lane claims any
                // keypoint of its list (8 shared-memory probes instead of one shuffle round per accepting lane)
                if (mine && st == 1) atomicMin(&claim[my_best_idx], lane);
                __syncwarp();
                if (mine) {
#pragma unroll
                    for (int k = 0; k < kTopK; ++k) conflict |= (keys[k] != kInfKey && claim[key_idx(keys[k])] < lane);
                }
                __syncwarp();
                if (mine && st == 1) claim[my_best_idx] = 255;
            } else {
                for (unsigned m = accept_mask; m; m &= m - 1) {
                    const int jl = __ffs(m) - 1;
                    const unsigned bj = __shfl_sync(0xFFFFFFFFu, my_best_idx, jl);
                    if (mine && lane > jl) {
#pragma unroll
                        for (int k = 0; k < kTopK; ++k) conflict |= (keys[k] != kInfKey && key_idx(keys[k]) == bj);
                    }
                }
            }
            const unsigned unsafe = __ballot_sync(0xFFFFFFFFu, mine && (conflict || st == 2));
            const int first_unsafe = unsafe ? __ffs(unsafe) - 1 : 32;
            const unsigned commit = pending & ((first_unsafe >= 32) ? 0xFFFFFFFFu : ((1u << first_unsafe) - 1u));
            if (((commit >> lane) & 1u) && st == 1) {
                const unsigned i1 = key_idx(best);
                atomicOr(&taken[i1 >> 5], 1u << (i1 & 31));  // distinct lanes may share a word
                m21[i1] = r;
            }
            pending &= ~commit;
            __syncwarp();
            if (first_unsafe < 32 && ((exact_mask >> first_unsafe) & 1u) && !(((unsafe & ((1u << first_unsafe) - 1u)) != 0u))) {
                // the first not-yet-committed row cannot be decided from its list: warp-cooperative exact scan (the reference's inner loop)
                const int rr = base + first_unsafe;
                unsigned bk, sd;
                exact_row(d1, a1p, a1s, n1, taken, desc2[(size_t)(b2 + rr) * 2], desc2[(size_t)(b2 + rr) * 2 + 1], side_angle(S2, b2 + rr),
                          check_orientation, lane, &bk, &sd);
                if (bk != kInfKey) {
                    const unsigned bd = key_dist(bk);
                    if (!(bd > (unsigned)kThrLow) && !(__fmul_rn(lowe_ratio, (float)sd) < (float)bd)) {
                        const unsigned i1 = key_idx(bk);
                        if (lane == 0) {
                            taken[i1 >> 5] |= 1u << (i1 & 31);
                            m21[i1] = rr;
                        }
                    }
                }
                pending &= ~(1u << first_unsafe);
                __syncwarp();
            }
        }
    }
    __syncwarp();
    // robust.cc:317-325: pairs sorted by idx_1
    int total = 0;
    for (int base = 0; base < n1; base += 32) {
        const int i = base + lane;
        const int v = (i < n1) ? m21[i] : -1;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, v >= 0);
        if (v >= 0) {
            const int pos = total + __popc(bal & ((1u << lane) - 1u));
            pairs[2 * (size_t)pos] = i;
            pairs[2 * (size_t)pos + 1] = v;
        }
        total += __popc(bal);
    }
    if (lane == 0) n_pairs[p] = total;
}

// ---------------------------------------------------------------------------------------------------------------
// Grid-guided projection matchers
//   match::projection::match_frame_and_landmarks      src/stella_vslam/match/projection.cc:13-93    (mode 0)
//   match::projection::match_current_and_last_frames  src/stella_vslam/match/projection.cc:95-207   (mode 1)
//   data::assign_keypoints_to_grid / get_keypoints_in_cell   src/stella_vslam/data/common.cc:83-190, data/common.h:60-68
// G1 builds the keypoint grid (cell-x major, indices ascending inside a cell = the reference's iteration order), G2 lets one
// thread per landmark enumerate its search window in that order and record (distance, octave, index) for every candidate
// that passes the static gates, G3 replays the reference's sequential loop (a match removes the keypoint from all later
// searches) with the same prefix-commit scheme as the brute-force resolve.
// ---------------------------------------------------------------------------------------------------------------
struct GuidedDev {
    int n_train, n_queries, grid_cols, grid_rows, cap;
    float min_x, max_x, min_y, max_y;
    const float *t_x, *t_y, *t_angle, *t_x_right;
    const unsigned char* t_octave;
    const uint4* t_desc;
    const uint4* q_desc;
    const float *q_x, *q_y, *q_margin, *q_x_right, *q_angle;
    const signed char *q_min_level, *q_max_level;
    const unsigned char* q_valid;
    const double* q_reproj;           // mode 3
    const float* inv_level_sigma_sq;  // mode 3, 256 entries
    int do_reproj;
    // scratch + outputs of this problem
    int *cell_start, *cell_items, *cell_cursor, *owner;
    uint2* lists;
    int* list_len;
    unsigned char* occupied;
    int* match_out;
    int* n_matches;
};

__device__ __forceinline__ int cell_index(const GuidedDev& g, float x, float y, double inv_w, double inv_h) {
    const int cx = __double2int_rd((double)__fsub_rn(x, g.min_x) * inv_w), cy = __double2int_rd((double)__fsub_rn(y, g.min_y) * inv_h);
    return (0 <= cx && cx < g.grid_cols && 0 <= cy && cy < g.grid_rows) ? cx * g.grid_rows + cy : -1;
}

// G1: one CTA.  cell_start[c .. c+1) delimits the ascending keypoint indices of cell c (cell = cx * rows + cy).
__global__ void __launch_bounds__(1024) guided_grid_kernel(const GuidedDev* __restrict__ gs) {
    const GuidedDev g = gs[blockIdx.x];
    int *cell_start = g.cell_start, *cell_items = g.cell_items, *cell_cursor = g.cell_cursor;
    const int n_cells = g.grid_cols * g.grid_rows;
    const double inv_w = (double)g.grid_cols / (double)__fsub_rn(g.max_x, g.min_x), inv_h = (double)g.grid_rows / (double)__fsub_rn(g.max_y, g.min_y);
    for (int c = threadIdx.x; c <= n_cells; c += blockDim.x) cell_start[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < g.n_train; i += blockDim.x) {
        const int c = cell_index(g, g.t_x[i], g.t_y[i], inv_w, inv_h);
        if (c >= 0) atomicAdd(&cell_start[c + 1], 1);
    }
    __syncthreads();
    {  // running sum over the n_cells + 1 counters (block-wide: a contiguous chunk per thread, shuffle scan of the chunk sums).  A single
       // thread walking the 3 073 cells of the 64 x 48 grid was most of this kernel's 150 us.
        __shared__ int warp_tot[32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int per = (n_cells + 1 + (int)blockDim.x - 1) / (int)blockDim.x;
        const int beg = min((int)threadIdx.x * per, n_cells + 1), end = min(beg + per, n_cells + 1);
        int sum = 0;
        for (int c = beg; c < end; ++c) sum += cell_start[c];
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, w, off);
                if (lane >= off) w += v;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        int run = incl - sum + (warp > 0 ? warp_tot[warp - 1] : 0);
        for (int c = beg; c < end; ++c) {
            run += cell_start[c];
            cell_start[c] = run;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n_cells; c += blockDim.x) cell_cursor[c] = cell_start[c];
    __syncthreads();
    for (int i = threadIdx.x; i < g.n_train; i += blockDim.x) {
        const int c = cell_index(g, g.t_x[i], g.t_y[i], inv_w, inv_h);
        if (c >= 0) cell_items[atomicAdd(&cell_cursor[c], 1)] = i;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n_cells; c += blockDim.x) {  // restore ascending index order inside each (short) cell list
        const int a = cell_start[c], b = cell_start[c + 1];
        for (int i = a + 1; i < b; ++i) {
            const int v = cell_items[i];
            int j = i - 1;
            while (j >= a && cell_items[j] > v) {
                cell_items[j + 1] = cell_items[j];
                --j;
            }
            cell_items[j + 1] = v;
        }
    }
}

// G2: one thread per landmark: enumerate the window, apply the static gates, record candidates in iteration order.
// entry = distance << 8 | octave, idx
__global__ void __launch_bounds__(128) guided_candidates_kernel(const GuidedDev* __restrict__ gs, int mode, int check_orientation,
                                                                int* __restrict__ overflow) {
    const GuidedDev g = gs[blockIdx.y];
    const int *cell_start = g.cell_start, *cell_items = g.cell_items;
    uint2* lists = g.lists;
    int* list_len = g.list_len;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.n_queries) return;
    int len = 0;
    if (!g.q_valid || (g.q_valid[q] & 1)) {
        const double inv_w = (double)g.grid_cols / (double)__fsub_rn(g.max_x, g.min_x), inv_h = (double)g.grid_rows / (double)__fsub_rn(g.max_y, g.min_y);
        const float ref_x = g.q_x[q], ref_y = g.q_y[q], margin = g.q_margin[q];
        const int min_level = g.q_min_level[q], max_level = g.q_max_level[q];
        // data/common.cc:137-155
        const int min_cx = max(0, __double2int_rd((double)__fsub_rn(__fsub_rn(ref_x, g.min_x), margin) * inv_w));
        const int max_cx = min(g.grid_cols - 1, __double2int_ru((double)__fadd_rn(__fsub_rn(ref_x, g.min_x), margin) * inv_w));
        const int min_cy = max(0, __double2int_rd((double)__fsub_rn(__fsub_rn(ref_y, g.min_y), margin) * inv_h));
        const int max_cy = min(g.grid_rows - 1, __double2int_ru((double)__fadd_rn(__fsub_rn(ref_y, g.min_y), margin) * inv_h));
        if (min_cx < g.grid_cols && max_cx >= 0 && min_cy < g.grid_rows && max_cy >= 0) {
            const uint4 q0 = g.q_desc[(size_t)q * 2], q1 = g.q_desc[(size_t)q * 2 + 1];
            uint2* out = lists + (size_t)q * g.cap;
            for (int cx = min_cx; cx <= max_cx; ++cx)
                for (int cy = min_cy; cy <= max_cy; ++cy) {
                    const int c = cx * g.grid_rows + cy;
                    for (int k = cell_start[c]; k < cell_start[c + 1]; ++k) {
                        const int idx = cell_items[k];
                        const int oct = g.t_octave[idx];
                        if (0 <= min_level && oct < min_level) continue;
                        if (0 <= max_level && max_level < oct) continue;
                        const float dx = __fsub_rn(g.t_x[idx], ref_x), dy = __fsub_rn(g.t_y[idx], ref_y);
                        if (!(fabsf(dx) < margin && fabsf(dy) < margin)) continue;
                        if (mode <= 1 && g.t_x_right) {  // stereo gate (projection.cc:56-61, 168-173)
                            const float xr = g.t_x_right[idx];
                            if (0.f < xr && margin < fabsf(__fsub_rn(g.q_x_right[q], xr))) continue;
                        }
                        if ((mode == 1 || mode == 4) && check_orientation && orientation_rejects(g.q_angle[q], g.t_angle[idx])) continue;
                        if (mode == 3 && g.do_reproj) {  // chi-square reprojection gate (fuse.cc:93-120), evaluated like the reference: double
                            const double e_x = __dsub_rn(g.q_reproj[2 * q], (double)g.t_x[idx]), e_y = __dsub_rn(g.q_reproj[2 * q + 1], (double)g.t_y[idx]);
                            double err_sq = __dadd_rn(__dmul_rn(e_x, e_x), __dmul_rn(e_y, e_y));
                            float chi_sq = 5.99146f;
                            if (g.t_x_right && g.t_x_right[idx] >= 0.f) {
                                const float e_xr = __fsub_rn(g.q_x_right[q], g.t_x_right[idx]);
                                err_sq = __dadd_rn(err_sq, (double)__fmul_rn(e_xr, e_xr));
                                chi_sq = 7.81473f;
                            }
                            if ((double)chi_sq < __dmul_rn(err_sq, (double)g.inv_level_sigma_sq[oct])) continue;
                        }
                        const unsigned d = hamming256(q0, q1, g.t_desc[(size_t)idx * 2], g.t_desc[(size_t)idx * 2 + 1]);
                        if (len < g.cap) out[len] = make_uint2((d << 8) | (unsigned)oct, (unsigned)idx);
                        ++len;
                    }
                }
        }
    }
    if (len > g.cap) {
        atomicMax(overflow, len);
        len = g.cap;
    }
    list_len[q] = len;
}

// the reference's best / second update in iteration order over the keypoints whose state admits the candidate
// (state: 0 = occupied, 0xFFFF = free; mode 4: Hamming distance of the match the keypoint currently holds, area.cc:49-51).
// Returns the accepted index or -1; *best_out = its distance.
__device__ __forceinline__ int guided_decide(const uint2* __restrict__ list, int len, const volatile unsigned short* state, int mode, unsigned thr,
                                             float lowe_ratio, unsigned* best_out) {
    // mode 5: bow_tree::match_frame_and_keyframe / match_keyframes; mode 6: match_for_triangulation, whose running best starts at
    // the threshold and which skips every candidate above it (robust.cc:57-59, 83-85)
    unsigned best = mode == 6 ? thr : (unsigned)kMaxDist, second = kMaxDist;
    int best_level = -1, second_level = -1, best_idx = -1;
    const bool track_second = mode == 0 || mode >= 4;
    for (int k = 0; k < len; ++k) {
        const uint2 e = list[k];
        const unsigned idx = e.y, d = e.x >> 8;
        if ((unsigned)state[idx] <= d) continue;
        if (mode == 6 && d > best) continue;
        const int oct = (int)(e.x & 0xFF);
        if (d < best) {
            second = best;
            second_level = best_level;
            best = d;
            best_level = oct;
            best_idx = (int)idx;
        } else if (track_second && d < second) {
            second_level = oct;
            second = d;
        }
    }
    *best_out = best;
    if (best_idx < 0 || best > thr) return -1;
    if (mode == 0 && best_level == second_level && (float)best > __fmul_rn(lowe_ratio, (float)second)) return -1;
    if (mode >= 4 && __fmul_rn((float)second, lowe_ratio) < (float)best) return -1;
    return best_idx;
}

// G3: one warp; shared memory: [claim per keypoint (int)][state per keypoint (u16)].  A lane's decision is safe to commit when
// no lower lane of the batch wants ANY keypoint of its list (that is the only way an earlier landmark can change a later one's
// outcome).  Mode 2 is stateless: every decision commits at once.
template <class Dev>
__global__ void __launch_bounds__(32) guided_resolve_kernel(const Dev* __restrict__ gs, int mode, unsigned thr, float lowe_ratio) {
    extern __shared__ unsigned guided_smem[];
    const Dev g = gs[blockIdx.x];
    const uint2* lists = g.lists;
    const int* list_len = g.list_len;
    unsigned char* occupied_io = g.occupied;
    int *match_out = g.match_out, *n_matches = g.n_matches, *owner = g.owner;
    const int lane = threadIdx.x;
    int* claim = reinterpret_cast<int*>(guided_smem);
    unsigned short* state = reinterpret_cast<unsigned short*>(guided_smem + g.n_train);
    for (int i = lane; i < g.n_train; i += 32) {
        claim[i] = 255;
        state[i] = mode == 4 ? (unsigned short)kMaxDist : ((occupied_io && occupied_io[i]) ? 0 : 0xFFFF);
        if (mode == 4) owner[i] = -1;
    }
    for (int q = lane; q < g.n_queries; q += 32) match_out[q] = -1;
    __syncwarp();
    int total = 0;
    for (int base = 0; base < g.n_queries; base += 32) {
        const int q = base + lane;
        const int len = (q < g.n_queries) ? list_len[q] : 0;
        const uint2* list = lists + (size_t)min(q, g.n_queries - 1) * g.cap;
        unsigned pending = __ballot_sync(0xFFFFFFFFu, len > 0);
        while (pending) {
            const bool mine = (pending >> lane) & 1u;
            unsigned best = kMaxDist;
            const int acc = mine ? guided_decide(list, len, state, mode, thr, lowe_ratio, &best) : -1;
            unsigned commit = pending;
            if (mode != 2) {
                if (acc >= 0) atomicMin(&claim[acc], lane);  // claim[idx] = lowest lane that wants idx this round
                __syncwarp();
                bool conflict = false;
                if (mine)
                    for (int k = 0; k < len; ++k) conflict |= claim[list[k].y] < lane;
                const unsigned unsafe = __ballot_sync(0xFFFFFFFFu, mine && conflict);
                if (unsafe) commit = pending & ((1u << (__ffs(unsafe) - 1)) - 1u);
                __syncwarp();
                if (acc >= 0) claim[acc] = 255;  // reset for the next round
            }
            const bool do_commit = ((commit >> lane) & 1u) && acc >= 0;
            int stolen = 0;
            if (do_commit) {
                match_out[q] = acc;
                if (mode == 4) {  // area.cc:75-87: take the keypoint over from its previous owner
                    const int prev = owner[acc];
                    if (prev >= 0) {
                        match_out[prev] = -1;
                        stolen = 1;
                    }
                    owner[acc] = q;
                    state[acc] = (unsigned short)best;
                } else if (mode != 2) {
                    // projection.cc:50-53, 163-166: a keypoint is closed to later landmarks only while the landmark it carries
                    // has_observation(); a temporal landmark (no observation yet) can be overwritten by a later one
                    if (!g.q_valid || (g.q_valid[q] & 2)) state[acc] = 0;
                }
            }
            total += __popc(__ballot_sync(0xFFFFFFFFu, do_commit)) - __popc(__ballot_sync(0xFFFFFFFFu, stolen));
            pending &= ~commit;
            __syncwarp();
        }
    }
    __syncwarp();
    if (mode == 0 || mode == 1 || mode == 3)
        for (int i = lane; i < g.n_train; i += 32) occupied_io[i] = state[i] == 0;
    if (lane == 0) *n_matches = total;
}


// b200_track_local_map: the keypoint count of a frame is known on the device only (the extractor's counter)
__global__ void track_set_counts_kernel(GuidedDev* __restrict__ gs, const chain::TrackFrameDev* __restrict__ frames, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) gs[f].n_train = frames[f].status[0];
}

// ---------------------------------------------------------------------------------------------------------------
// All-pairs matchers with greedy state other than brute_force_match:
//   match::bow_tree::match_frame_and_keyframe   src/stella_vslam/match/bow_tree.cc:169-256   (variant 0)
//   match::bow_tree::match_keyframes            bow_tree.cc:258-366                          (variant 0)
//   match::robust::match_for_triangulation      src/stella_vslam/match/robust.cc:14-146      (variant 1)
//   match::bow_tree::match_for_triangulation    bow_tree.cc:11-167                           (variant 1, with node ids)
// P1 evaluates every (row, candidate) pair once and records, per row and in candidate order, the candidates that pass the
// state-independent gates with a distance that can still influence the outcome (<= list_thr); the sequential part is then the
// same warp-batched replay as for the guided matchers (guided_resolve_kernel, modes 5 and 6).  A BoW node holds each keypoint
// exactly once and rows of different nodes never compete for a candidate, so visiting rows in index order with the gate
// node_1[i] == node_2[j] gives the merge-join's result.
// ---------------------------------------------------------------------------------------------------------------
struct PairsDev {
    int n_queries, n_train, cap;  // rows (side 1), candidates (side 2)
    const uint4 *desc1, *desc2;
    const float *angle1, *angle2;
    const unsigned char *valid1, *valid2, *stereo1, *stereo2;
    const int *node1, *node2;
    const double *bearing1, *bearing2;
    const float* scale1;
    double E[9], epi[3];
    int valid_epiplane;
    float residual_rad_thr;
    uint2* lists;
    int* list_len;
    unsigned char* occupied;  // always null: every candidate starts free
    const unsigned char* q_valid;  // always null (interface of the shared resolve kernel)
    int *match_out, *n_matches, *owner;
};

// match/base.h:67-79 in the reference's evaluation order (3x3 times 3, dot, norm; no contraction)
__device__ __forceinline__ bool epipolar_inlier(const double* __restrict__ b1, const double* __restrict__ b2, const double* E, float thr, float scale) {
    const double x = b2[0], y = b2[1], z = b2[2];
    const double e0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x), __dmul_rn(E[1], y)), __dmul_rn(E[2], z));
    const double e1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], x), __dmul_rn(E[4], y)), __dmul_rn(E[5], z));
    const double e2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], x), __dmul_rn(E[7], y)), __dmul_rn(E[8], z));
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(e0, b1[0]), __dmul_rn(e1, b1[1])), __dmul_rn(e2, b1[2]));
    const double norm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(e2, e2)));
    double c = __ddiv_rn(dot, norm);
    c = fmax(-1.0, c);
    c = fmin(1.0, c);
    const double residual_rad = fabs(__dsub_rn(1.5707963267948966, acos(c)));
    return residual_rad < (double)__fmul_rn(thr, scale);
}

constexpr int kPairRows = 64, kPairChunk = 256;

__global__ void __launch_bounds__(kPairRows) pairs_candidates_kernel(const PairsDev* __restrict__ ps, int variant, unsigned list_thr,
                                                                    int check_orientation, int* __restrict__ overflow) {
    __shared__ uint4 s2[kPairChunk * 2];
    __shared__ float sa[kPairChunk];
    __shared__ int sn[kPairChunk];
    __shared__ unsigned char sv[kPairChunk];
    const PairsDev& g = ps[blockIdx.y];
    const int n1 = g.n_queries, n2 = g.n_train;
    if ((int)(blockIdx.x * kPairRows) >= n1) return;
    const int row = blockIdx.x * kPairRows + threadIdx.x;
    const bool active = row < n1 && (!g.valid1 || g.valid1[row]);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    float qa = 0.f;
    int qn = 0;
    bool q_stereo = false;
    if (active) {
        q0 = g.desc1[(size_t)row * 2];
        q1 = g.desc1[(size_t)row * 2 + 1];
        if (check_orientation) qa = g.angle1[row];
        if (g.node1) qn = g.node1[row];
        q_stereo = g.stereo1 && g.stereo1[row];
    }
    uint2* out = g.lists + (size_t)min(row, n1 - 1) * g.cap;
    int len = 0;
    for (int c0 = 0; c0 < n2; c0 += kPairChunk) {
        const int cn = min(kPairChunk, n2 - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < cn * 2; t += blockDim.x) s2[t] = g.desc2[(size_t)c0 * 2 + t];
        for (int t = threadIdx.x; t < cn; t += blockDim.x) {
            sa[t] = check_orientation ? g.angle2[c0 + t] : 0.f;
            sn[t] = g.node2 ? g.node2[c0 + t] : 0;
            sv[t] = g.valid2 ? g.valid2[c0 + t] : 1;
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int j = 0; j < cn; ++j) {
            const unsigned dist = hamming256(q0, q1, s2[2 * j], s2[2 * j + 1]);
            if (dist > list_thr) continue;
            if (!sv[j] || sn[j] != qn) continue;
            if (check_orientation && orientation_rejects(qa, sa[j])) continue;
            if (variant == 1) {
                const double* b2 = g.bearing2 + (size_t)(c0 + j) * 3;
                if (g.valid_epiplane && !q_stereo && !(g.stereo2 && g.stereo2[c0 + j])) {  // robust.cc:87-98: too close to the epipole
                    const double cos_dist = __dadd_rn(__dadd_rn(__dmul_rn(g.epi[0], b2[0]), __dmul_rn(g.epi[1], b2[1])), __dmul_rn(g.epi[2], b2[2]));
                    if (0.99862953475 < cos_dist) continue;
                }
                if (!epipolar_inlier(g.bearing1 + (size_t)row * 3, b2, g.E, g.residual_rad_thr, g.scale1[row])) continue;
            }
            if (len < g.cap) out[len] = make_uint2(dist << 8, (unsigned)(c0 + j));
            ++len;
        }
    }
    if (row < n1) {
        if (len > g.cap) {
            atomicMax(overflow, len);
            len = g.cap;
        }
        g.list_len[row] = len;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// match::stereo  (src/stella_vslam/match/stereo.cc:20-251): row-band candidates, arg-min Hamming < 75, 11x11 L1 patch correlation
// over +-5 px on the keypoint's pyramid level, parabola sub-pixel, rejection above twice the median correlation.
// Left keypoints are independent of each other: S1 (thread per left keypoint, right keypoints streamed through shared memory),
// S2 (warp per left keypoint: patch correlation on the pyramid levels that the two extractors keep on the device),
// S3 (one CTA: exact median by two-pass radix select, then the rejection).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStereoMaxLevels = 16;
constexpr unsigned kStereoThr = (100u + 50u) / 2u;  // stereo.h:99
struct StereoLevel {
    const unsigned char *left, *right;
    unsigned long long pitch_l, pitch_r;
    int w, h;
    float sf, inv_sf;
};
struct StereoDev {
    StereoLevel lv[kStereoMaxLevels];
    int n_levels, n_left, n_right;
    const b200_keypoint_t *kl, *kr;
    const uint4 *dl, *dr;
    float fxb, max_disp;
    int* best_right;  // [n_left]
    float *x_right, *depth;
    int* corr;        // [n_left], -1 = no stereo match
    int* n_kept;
};

constexpr int kStereoRows = 64, kStereoChunk = 256;

__global__ void __launch_bounds__(kStereoRows) stereo_match_kernel(const StereoDev* __restrict__ sp) {
    __shared__ uint4 sd[kStereoChunk * 2];
    __shared__ float sx[kStereoChunk];
    __shared__ int slo[kStereoChunk], shi[kStereoChunk], soct[kStereoChunk];
    const StereoDev& g = *sp;
    const int i = blockIdx.x * kStereoRows + threadIdx.x;
    const bool active = i < g.n_left;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    int row = -1, level = 0;
    float min_x = 0.f, max_x = -1.f;
    if (active) {
        const b200_keypoint_t k = g.kl[i];
        q0 = g.dl[(size_t)i * 2];
        q1 = g.dl[(size_t)i * 2 + 1];
        level = k.octave;
        row = (int)k.y;                       // indices_right_in_row.at(y_left): float -> size_t truncation (stereo.cc:41)
        min_x = __fsub_rn(k.x, g.max_disp);   // stereo.cc:47-48 (min_disp_ = 0)
        max_x = k.x;
    }
    const bool searching = active && !(max_x < 0.f);
    unsigned best = make_key(kStereoThr, 0);  // only strictly smaller distances win (stereo.cc:172)
    for (int c0 = 0; c0 < g.n_right; c0 += kStereoChunk) {
        const int cn = min(kStereoChunk, g.n_right - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < cn * 2; t += blockDim.x) sd[t] = g.dr[(size_t)c0 * 2 + t];
        for (int t = threadIdx.x; t < cn; t += blockDim.x) {
            const b200_keypoint_t k = g.kr[c0 + t];
            const float r = __fmul_rn(2.0f, g.lv[min(max(k.octave, 0), g.n_levels - 1)].sf);  // stereo.cc:131-135
            sx[t] = k.x;
            shi[t] = __float2int_ru(__fadd_rn(k.y, r));
            slo[t] = __float2int_rd(__fsub_rn(k.y, r));
            soct[t] = k.octave;
        }
        __syncthreads();
        if (!searching) continue;
        for (int j = 0; j < cn; ++j) {
            if (row < slo[j] || shi[j] < row) continue;
            if (soct[j] < level - 1 || soct[j] > level + 1) continue;  // stereo.cc:158-160
            if (sx[j] < min_x || max_x < sx[j]) continue;              // stereo.cc:163-166
            const unsigned key = make_key(hamming256(q0, q1, sd[2 * j], sd[2 * j + 1]), (unsigned)(c0 + j));
            if (key_dist(key) < key_dist(best)) best = key;           // first minimum in index order
        }
    }
    if (active) g.best_right[i] = (key_dist(best) < kStereoThr) ? (int)key_idx(best) : -1;
}

__global__ void __launch_bounds__(128) stereo_subpixel_kernel(const StereoDev* __restrict__ sp) {
    const StereoDev& g = *sp;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= g.n_left) return;
    const int j = g.best_right[i];
    float out_x = -1.f, out_depth = -1.f;
    int out_corr = -1;
    if (j >= 0) {
        const b200_keypoint_t kl = g.kl[i];
        const StereoLevel& L = g.lv[min(max(kl.octave, 0), g.n_levels - 1)];
        const float x_right = g.kr[j].x;
        const int sxl = __float2int_rn(__fmul_rn(kl.x, L.inv_sf)), syl = __float2int_rn(__fmul_rn(kl.y, L.inv_sf));
        const int sxr = __float2int_rn(__fmul_rn(x_right, L.inv_sf));
        constexpr int win = 5, slide = 5;
        const bool in_range = !(sxr - slide - win < 0 || L.w <= sxr + slide + win);  // stereo.cc:193-197
        // the reference's rowRange/colRange would assert outside the image; the extractor's 19-px border keeps patches inside
        const bool patch_ok = sxl - win >= 0 && sxl + win < L.w && syl - win >= 0 && syl + win < L.h;
        if (in_range && patch_ok) {
            const unsigned char* pl = L.left + (size_t)syl * L.pitch_l + sxl;
            const unsigned char* pr = L.right + (size_t)syl * L.pitch_r + sxr;
            const int lc = pl[0];
            int lv[4], dyv[4], dxv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = min(lane + 32 * u, 120);
                dyv[u] = p / 11 - win;
                dxv[u] = p % 11 - win;
                lv[u] = (int)pl[(long long)dyv[u] * (long long)L.pitch_l + dxv[u]] - lc;
            }
            int best_corr = 0x7FFFFFFF, best_offset = 0, c_prev = 0, c1 = 0, c2 = 0, c3 = 0;
            bool want_next = false;
            for (int off = -slide; off <= slide; ++off) {
                const int rc = pr[off];
                int s = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (lane + 32 * u < 121) s += abs(lv[u] - ((int)pr[(long long)dyv[u] * (long long)L.pitch_r + dxv[u] + off] - rc));
                s = __reduce_add_sync(0xFFFFFFFFu, s);
                if (want_next) {
                    c3 = s;
                    want_next = false;
                }
                if (s < best_corr) {  // strict: the first minimum wins (stereo.cc:221-224)
                    best_corr = s;
                    best_offset = off;
                    c1 = c_prev;
                    c2 = s;
                    want_next = true;
                }
                c_prev = s;
            }
            if (best_offset != -slide && best_offset != slide) {
                const float f1 = (float)c1, f2 = (float)c2, f3 = (float)c3;
                const double num = (double)__fsub_rn(f1, f3);
                const double den = __dsub_rn(__dmul_rn(2.0, (double)__fadd_rn(f1, f3)), __dmul_rn(4.0, (double)f2));
                const float x_delta = __double2float_rn(__ddiv_rn(num, den));
                if (!(x_delta < -1.0f || 1.0f < x_delta)) {
                    float best_x = __fmul_rn(L.sf, __fadd_rn((float)(sxr + best_offset), x_delta));
                    float disp = __fsub_rn(kl.x, best_x);
                    if (!(disp < 0.f || g.max_disp <= disp)) {
                        if (disp <= 0.f) {  // stereo.cc:78-82
                            disp = 0.01f;
                            best_x = __fsub_rn(kl.x, disp);
                        }
                        out_depth = __fdiv_rn(g.fxb, disp);
                        out_x = best_x;
                        out_corr = best_corr;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        g.x_right[i] = out_x;
        g.depth[i] = out_depth;
        g.corr[i] = out_corr;
    }
}

// exact median (element size/2 of the ascending order) of the valid correlations, then stereo.cc:96-113
__global__ void __launch_bounds__(1024) stereo_median_kernel(const StereoDev* __restrict__ sp) {
    __shared__ int hist[256];
    __shared__ int sel_hi, sel_rank, n_valid, median, n_rejected;
    const StereoDev& g = *sp;
    const int tid = threadIdx.x;
    if (tid < 256) hist[tid] = 0;
    if (tid == 0) n_valid = 0, n_rejected = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < g.n_left; i += blockDim.x) {
        const int c = g.corr[i];
        if (c >= 0) {
            atomicAdd(&hist[min(c >> 8, 255)], 1);  // correlations are <= 121 * 510 < 65536
            ++local;
        }
    }
    atomicAdd(&n_valid, local);
    __syncthreads();
    if (n_valid == 0) {
        if (tid == 0) *g.n_kept = 0;
        return;
    }
    if (tid == 0) {
        int k = n_valid / 2, b = 0;
        while (k >= hist[b]) k -= hist[b++];
        sel_hi = b;
        sel_rank = k;
    }
    __syncthreads();
    const int hi = sel_hi;
    __syncthreads();
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < g.n_left; i += blockDim.x) {
        const int c = g.corr[i];
        if (c >= 0 && min(c >> 8, 255) == hi) atomicAdd(&hist[c & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int k = sel_rank, b = 0;
        while (k >= hist[b]) k -= hist[b++];
        median = (hi << 8) | b;
    }
    __syncthreads();
    const float thr = __double2float_rn(__dmul_rn(2.0, (double)(float)median));
    int rejected = 0;
    for (int i = tid; i < g.n_left; i += blockDim.x) {
        const int c = g.corr[i];
        if (c >= 0 && thr < (float)c) {
            g.x_right[i] = -1.f;
            g.depth[i] = -1.f;
            ++rejected;
        }
    }
    atomicAdd(&n_rejected, rejected);
    __syncthreads();
    if (tid == 0) *g.n_kept = n_valid - n_rejected;
}

// ---------------------------------------------------------------------------------------------------------------
// data::landmark::compute_descriptor (src/stella_vslam/data/landmark.cc:199-256): the representative descriptor of a
// landmark = the observation whose median Hamming distance to all observations is smallest (first index on ties).  One warp per
// landmark; row by row the lanes compute the distances into shared memory and select the element [0.5 (n - 1)] of the sorted row
// by counting (value v is the k-th smallest iff #(d < v) <= k < #(d <= v)).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLmMaxObs = 512;  // observations of one landmark handled on chip
__global__ void __launch_bounds__(128) landmark_descriptor_kernel(const uint4* __restrict__ descs, const int* __restrict__ offsets, int n_landmarks,
                                                                  int* __restrict__ best_idx_out, uint4* __restrict__ desc_out) {
    __shared__ unsigned short dist[4][kLmMaxObs];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * 4 + warp;
    if (l >= n_landmarks) return;
    const int o = offsets[l], n = offsets[l + 1] - o;
    if (n <= 0 || n > kLmMaxObs) {
        if (lane == 0) best_idx_out[l] = n <= 0 ? -1 : -2;  // -2: more observations than the kernel holds (reported by the host)
        return;
    }
    const int k = (int)(0.5 * (n - 1));
    unsigned best = kMaxDist;
    int best_idx = 0;
    unsigned short* d = dist[warp];
    for (int i = 0; i < n; ++i) {
        const uint4 a0 = descs[(size_t)(o + i) * 2], a1 = descs[(size_t)(o + i) * 2 + 1];
        for (int j = lane; j < n; j += 32) d[j] = (unsigned short)hamming256(a0, a1, descs[(size_t)(o + j) * 2], descs[(size_t)(o + j) * 2 + 1]);
        __syncwarp();
        unsigned med = kMaxDist + 1;
        for (int j = lane; j < n; j += 32) {
            const unsigned v = d[j];
            int lt = 0, le = 0;
            for (int t = 0; t < n; ++t) {
                const unsigned w = d[t];
                lt += w < v;
                le += w <= v;
            }
            if (lt <= k && k < le) med = min(med, v);
        }
        med = warp_min(med);
        if (med < best) {
            best = med;
            best_idx = i;
        }
        __syncwarp();
    }
    if (lane == 0) best_idx_out[l] = best_idx;
    if (desc_out && lane < 2) desc_out[(size_t)l * 2 + lane] = descs[(size_t)(o + best_idx) * 2 + lane];
}

// data::landmark::update_mean_normal_and_obs_scale_variance (src/stella_vslam/data/landmark.cc:256-311): per landmark the
// normalised mean of the unit viewing directions of its observations and the ORB scale range from its reference keyframe.
// Thread per landmark; explicit round-to-nearest operations (this file is compiled with FMA contraction on).
__device__ __forceinline__ double norm3(double x, double y, double z) {
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}
__global__ void __launch_bounds__(128) landmark_geometry_kernel(int n, const double* __restrict__ pos_w, const int* __restrict__ offsets,
                                                                const double* __restrict__ cam_centers, const double* __restrict__ ref_center,
                                                                const float* __restrict__ ref_scale, float inv_scale_last,
                                                                double* __restrict__ mean_normal, float* __restrict__ max_valid,
                                                                float* __restrict__ min_valid) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const double px = pos_w[3 * (size_t)l], py = pos_w[3 * (size_t)l + 1], pz = pos_w[3 * (size_t)l + 2];
    double mx = 0.0, my = 0.0, mz = 0.0;
    for (int o = offsets[l]; o < offsets[l + 1]; ++o) {
        const double vx = __dsub_rn(px, cam_centers[3 * (size_t)o]), vy = __dsub_rn(py, cam_centers[3 * (size_t)o + 1]),
                     vz = __dsub_rn(pz, cam_centers[3 * (size_t)o + 2]);
        const double nrm = norm3(vx, vy, vz);
        const bool pos = nrm > 0.0;  // Eigen normalized(): unchanged when the norm is 0
        mx = __dadd_rn(mx, pos ? __ddiv_rn(vx, nrm) : vx);
        my = __dadd_rn(my, pos ? __ddiv_rn(vy, nrm) : vy);
        mz = __dadd_rn(mz, pos ? __ddiv_rn(vz, nrm) : vz);
    }
    const double mn = norm3(mx, my, mz);
    mean_normal[3 * (size_t)l] = mn > 0.0 ? __ddiv_rn(mx, mn) : mx;
    mean_normal[3 * (size_t)l + 1] = mn > 0.0 ? __ddiv_rn(my, mn) : my;
    mean_normal[3 * (size_t)l + 2] = mn > 0.0 ? __ddiv_rn(mz, mn) : mz;
    const double dist = norm3(__dsub_rn(px, ref_center[3 * (size_t)l]), __dsub_rn(py, ref_center[3 * (size_t)l + 1]), __dsub_rn(pz, ref_center[3 * (size_t)l + 2]));
    const float mxv = __double2float_rn(__dmul_rn(dist, (double)ref_scale[l]));
    max_valid[l] = mxv;
    min_valid[l] = __fmul_rn(mxv, inv_scale_last);
}

struct Matcher {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // b200_matcher_set_async_resolve: the sequential resolve pass (64 warps on the whole chip, ~0.5 ms per 64 pairs) runs on a side
    // stream so that the caller's next kernels (the next batch's extraction) fill the idle SMs; joined by the next matcher call
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_topk = nullptr, ev_resolved = nullptr;
    bool async_resolve = false, resolve_pending = false;
    bool timing = false;
    cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_track[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // b200_track_local_map stage boundaries
    bool track_timed = false;
    // all-pairs distances + top-K on the tensor cores (tcgen05, int8 +-1 GEMM) or on the POPC pipe (B200_MATCH_TOPK=popc)
    bool use_tensor_core = true;
    int join() {  // the main stream waits for the side stream's resolve
        if (resolve_pending) {
            B200_CUDA(cudaStreamWaitEvent(stream, ev_resolved, 0));
            resolve_pending = false;
        }
        return B200_OK;
    }
    // scratch (grown on demand)
    unsigned* d_lists = nullptr;
    size_t lists_cap = 0;
    int* d_matched = nullptr;
    size_t matched_cap = 0;
    unsigned* d_taken = nullptr;
    size_t taken_cap = 0;
    // staging for the host-buffer entry points
    unsigned char* d_stage = nullptr;
    size_t stage_cap = 0;
    size_t last_h2d = 0, last_d2h = 0;
    // grid-guided matchers: pinned host staging + device arena
    unsigned char* h_guided = nullptr;
    size_t h_guided_cap = 0;
    unsigned char* d_guided = nullptr;
    size_t d_guided_cap = 0;

    int grow_pinned(unsigned char** p, size_t* cap, size_t bytes) {
        if (bytes <= *cap) return B200_OK;
        B200_CUDA(cudaStreamSynchronize(stream));
        if (*p) B200_CUDA(cudaFreeHost(*p));
        *p = nullptr;
        *cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        B200_CUDA(cudaMallocHost((void**)p, want));
        *cap = want;
        return B200_OK;
    }

    int grow(void** p, size_t* cap, size_t bytes) {
        if (bytes <= *cap) return B200_OK;
        B200_CUDA(cudaStreamSynchronize(stream));
        if (*p) B200_CUDA(cudaFree(*p));
        *p = nullptr;
        *cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        B200_CUDA(cudaMalloc(p, want));
        *cap = want;
        return B200_OK;
    }

    int run(int n_problems, const Side& S1, const Side& S2, const void* valid2, int max_n1, int max_n2, float lowe, int check_ori,
            void* pairs, int pairs_stride, void* n_pairs, bool device_call = false) {
        if (n_problems <= 0) return B200_OK;
        if (max_n1 >= (1 << 22) - 1) {
            set_error("brute-force matcher supports < 4194304 keypoints per frame");
            return B200_ERR_INVALID;
        }
        int rc;
        if ((rc = join())) return rc;  // (the previous resolve still reads the candidate lists this call overwrites)
        max_n1 = std::max(max_n1, 1);
        if (pairs_stride < max_n1) {
            set_error("pairs_stride %d is smaller than the largest frame (%d keypoints)", pairs_stride, max_n1);
            return B200_ERR_CAPACITY;
        }
        const int taken_words = ceil_div(max_n1, 32);
        const int row_blocks = std::max(1, ceil_div(max_n2, kRowsPerBlock)), list_rows = row_blocks * kRowsPerBlock;
        if ((rc = grow((void**)&d_lists, &lists_cap, sizeof(unsigned) * kTopK * (size_t)list_rows * n_problems))) return rc;
        if ((rc = grow((void**)&d_matched, &matched_cap, sizeof(int) * (size_t)max_n1 * n_problems))) return rc;
        if ((rc = grow((void**)&d_taken, &taken_cap, sizeof(unsigned) * (size_t)taken_words * n_problems))) return rc;
        if (timing) B200_CUDA(cudaEventRecord(ev_t[0], stream));
        if (use_tensor_core) {
            const float need = lowe > 0.f ? std::ceil((float)kThrLow / lowe) + 1.f : (float)kMaxDist;
            const unsigned cap = (unsigned)std::min((float)kMaxDist, std::max((float)kThrLow, need));
            B200_CUDA(cudaFuncSetAttribute(topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem) + 1024));
            topk_tc_kernel<<<dim3(ceil_div(std::max(max_n2, 1), kTcRows), n_problems), kTcThreads, sizeof(TcSmem) + 1024, stream>>>(
                S1, S2, (const unsigned char*)valid2, check_ori, d_lists, list_rows, cap);
        } else {
            topk_kernel<<<dim3(row_blocks, n_problems), kRowsPerBlock, 0, stream>>>(S1, S2, (const unsigned char*)valid2, check_ori, d_lists);
        }
        if (timing) B200_CUDA(cudaEventRecord(ev_t[1], stream));
        const size_t state_bytes = sizeof(unsigned) * ((size_t)taken_words + 2 * (size_t)max_n1);  // taken bitmap, idx_1 -> idx_2 table, claim table
        const size_t stage_bytes = sizeof(unsigned) * 9 * (size_t)max_n1;
        const int use_smem = (state_bytes + stage_bytes <= 200 * 1024) ? 2 : (state_bytes <= 200 * 1024 ? 1 : 0);
        const size_t rs_bytes = use_smem == 2 ? state_bytes + stage_bytes : state_bytes;
        if (use_smem && rs_bytes > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_bytes));
        cudaStream_t rs = stream;
        if (async_resolve && device_call) {
            B200_CUDA(cudaEventRecord(ev_topk, stream));
            B200_CUDA(cudaStreamWaitEvent(side_stream, ev_topk, 0));
            rs = side_stream;
        }
        resolve_kernel<<<n_problems, 32, use_smem ? rs_bytes : 0, rs>>>(S1, S2, (const unsigned char*)valid2, d_lists, lowe, check_ori, d_matched,
                                                                        d_taken, taken_words, list_rows, max_n1, pairs_stride, (int*)pairs,
                                                                        (int*)n_pairs, use_smem);
        B200_CUDA(cudaGetLastError());
        if (timing) B200_CUDA(cudaEventRecord(ev_t[2], rs));
        if (rs != stream) {
            B200_CUDA(cudaEventRecord(ev_resolved, side_stream));
            resolve_pending = true;
        }
        return B200_OK;
    }
};

}  // namespace match
}  // namespace b200

struct b200_matcher_s {
    b200::match::Matcher m;
};

using b200::match::Side;

extern "C" {

int b200_matcher_create(int device, b200_matcher_t* out) {
    if (!out) return B200_ERR_INVALID;
    int rc = b200::require_device(device);
    if (rc) return rc;
    b200_matcher_s* h = new (std::nothrow) b200_matcher_s();
    if (!h) return B200_ERR_INVALID;
    h->m.device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->m.own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->m.side_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->m.ev_topk, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->m.ev_resolved, cudaEventDisableTiming);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->m.ev_t[i]);
    if (e != cudaSuccess) {
        delete h;
        return b200::cuda_fail(e, "stream creation", __FILE__, __LINE__);
    }
    h->m.stream = h->m.own_stream;
    if (const char* tk = getenv("B200_MATCH_TOPK")) h->m.use_tensor_core = !(tk[0] == 'p');
    *out = h;
    return B200_OK;
}

int b200_matcher_destroy(b200_matcher_t h) {
    if (!h) return B200_OK;
    cudaSetDevice(h->m.device);
    cudaStreamSynchronize(h->m.stream);
    if (h->m.side_stream) cudaStreamSynchronize(h->m.side_stream);
    cudaFree(h->m.d_lists);
    cudaFree(h->m.d_matched);
    cudaFree(h->m.d_taken);
    cudaFree(h->m.d_stage);
    cudaFree(h->m.d_guided);
    if (h->m.h_guided) cudaFreeHost(h->m.h_guided);
    for (int i = 0; i < 3; ++i)
        if (h->m.ev_t[i]) cudaEventDestroy(h->m.ev_t[i]);
    for (int i = 0; i < 7; ++i)
        if (h->m.ev_track[i]) cudaEventDestroy(h->m.ev_track[i]);
    if (h->m.ev_topk) cudaEventDestroy(h->m.ev_topk);
    if (h->m.ev_resolved) cudaEventDestroy(h->m.ev_resolved);
    if (h->m.side_stream) cudaStreamDestroy(h->m.side_stream);
    if (h->m.own_stream) cudaStreamDestroy(h->m.own_stream);
    delete h;
    return B200_OK;
}

int b200_matcher_set_async_resolve(b200_matcher_t h, int enable) {
    if (!h) return B200_ERR_INVALID;
    int rc = h->m.join();
    if (rc) return rc;
    h->m.async_resolve = enable != 0;
    return B200_OK;
}

int b200_matcher_enable_timing(b200_matcher_t h, int enable) {
    if (!h) return B200_ERR_INVALID;
    h->m.timing = enable != 0;
    return B200_OK;
}

int b200_matcher_stage_ms(b200_matcher_t h, int stage, float* ms) {
    if (!h || !ms || stage < 0 || stage > 1) return B200_ERR_INVALID;
    B200_CUDA(cudaEventSynchronize(h->m.ev_t[stage + 1]));
    B200_CUDA(cudaEventElapsedTime(ms, h->m.ev_t[stage], h->m.ev_t[stage + 1]));
    return B200_OK;
}

int b200_matcher_join(b200_matcher_t h) {
    if (!h) return B200_ERR_INVALID;
    return h->m.join();
}

int b200_matcher_set_stream(b200_matcher_t h, void* stream, int use_own) {
    if (!h) return B200_ERR_INVALID;
    B200_CUDA(cudaStreamSynchronize(h->m.stream));
    if (h->m.side_stream) B200_CUDA(cudaStreamSynchronize(h->m.side_stream));
    h->m.resolve_pending = false;
    h->m.stream = use_own ? h->m.own_stream : (cudaStream_t)stream;
    return B200_OK;
}

int b200_matcher_sync(b200_matcher_t h) {
    if (!h) return B200_ERR_INVALID;
    if (h->m.resolve_pending) {
        B200_CUDA(cudaStreamSynchronize(h->m.side_stream));
        h->m.resolve_pending = false;
    }
    B200_CUDA(cudaStreamSynchronize(h->m.stream));
    return B200_OK;
}

int b200_match_bruteforce_device(b200_matcher_t h, int n_problems, const void* d_desc1, const void* d_angle1, size_t angle1_stride,
                                 const void* d_off1, const void* d_cnt1, const void* d_desc2, const void* d_angle2, size_t angle2_stride,
                                 const void* d_valid2, const void* d_off2, const void* d_cnt2, int max_n1, int max_n2, float lowe_ratio,
                                 int check_orientation, void* d_pairs, int pairs_stride, void* d_n_pairs) {
    B200_RANGE("b200:match:bruteforce_device");
    if (!h || n_problems < 0 || max_n1 < 0 || max_n2 < 0) return B200_ERR_INVALID;
    if (n_problems == 0) return B200_OK;
    if (!d_off1 || !d_off2 || !d_cnt1 || !d_cnt2 || !d_pairs || !d_n_pairs || !d_desc1 || !d_angle1 || !d_desc2 || !d_angle2) {
        b200::set_error("b200_match_bruteforce_device: null argument");
        return B200_ERR_INVALID;
    }
    B200_CUDA(cudaSetDevice(h->m.device));
    const Side S1{(const uint4*)d_desc1, (const unsigned char*)d_angle1, (long long)angle1_stride, (const int*)d_off1, (const int*)d_cnt1};
    const Side S2{(const uint4*)d_desc2, (const unsigned char*)d_angle2, (long long)angle2_stride, (const int*)d_off2, (const int*)d_cnt2};
    return h->m.run(n_problems, S1, S2, d_valid2, max_n1, max_n2, lowe_ratio, check_orientation, d_pairs, pairs_stride, d_n_pairs, true);
}

int b200_match_bruteforce(b200_matcher_t h, int n_problems, const uint8_t* desc1, const void* angle1, size_t angle1_stride,
                          const int32_t* off1, const int32_t* cnt1, const uint8_t* desc2, const void* angle2, size_t angle2_stride,
                          const uint8_t* valid2, const int32_t* off2, const int32_t* cnt2, float lowe_ratio, int check_orientation,
                          int32_t* pairs, int pairs_stride, int32_t* n_pairs) {
    B200_RANGE("b200:match:bruteforce");
    if (!h || n_problems < 0) return B200_ERR_INVALID;
    if (n_problems == 0) return B200_OK;
    if (!off1 || !off2 || !cnt1 || !cnt2 || !pairs || !n_pairs || angle1_stride < sizeof(float) || angle2_stride < sizeof(float)) {
        b200::set_error("b200_match_bruteforce: null argument or angle stride < 4");
        return B200_ERR_INVALID;
    }
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    // extents of the two sides that the problems touch
    int lo1 = INT_MAX, hi1 = 0, lo2 = INT_MAX, hi2 = 0, max_n1 = 0, max_n2 = 0;
    for (int p = 0; p < n_problems; ++p) {
        if (off1[p] < 0 || off2[p] < 0 || cnt1[p] < 0 || cnt2[p] < 0) {
            b200::set_error("b200_match_bruteforce: negative offset/count in problem %d", p);
            return B200_ERR_INVALID;
        }
        if (cnt1[p] > 0) { lo1 = std::min(lo1, off1[p]); hi1 = std::max(hi1, off1[p] + cnt1[p]); }
        if (cnt2[p] > 0) { lo2 = std::min(lo2, off2[p]); hi2 = std::max(hi2, off2[p] + cnt2[p]); }
        max_n1 = std::max(max_n1, cnt1[p]);
        max_n2 = std::max(max_n2, cnt2[p]);
    }
    if (hi1 == 0) lo1 = 0;
    if (hi2 == 0) lo2 = 0;
    const int ext1 = hi1 - lo1, ext2 = hi2 - lo2;
    if ((ext1 > 0 && (!desc1 || !angle1)) || (ext2 > 0 && (!desc2 || !angle2))) {
        b200::set_error("b200_match_bruteforce: null descriptor/angle buffer");
        return B200_ERR_INVALID;
    }
    if (pairs_stride < std::max(max_n1, 1)) {
        b200::set_error("pairs_stride %d is smaller than the largest frame (%d keypoints)", pairs_stride, max_n1);
        return B200_ERR_CAPACITY;
    }
    std::vector<int> meta(4 * (size_t)n_problems);  // off1 | cnt1 | off2 | cnt2, rebased to the staged extents
    for (int p = 0; p < n_problems; ++p) {
        meta[p] = cnt1[p] > 0 ? off1[p] - lo1 : 0;
        meta[n_problems + p] = cnt1[p];
        meta[2 * n_problems + p] = cnt2[p] > 0 ? off2[p] - lo2 : 0;
        meta[3 * n_problems + p] = cnt2[p];
    }
    // device staging (256-byte aligned): desc1 | desc2 | angle1 | angle2 | valid2 | meta | pairs | n_pairs
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    const size_t a1_bytes = ext1 > 0 ? (size_t)(ext1 - 1) * angle1_stride + sizeof(float) : 0;
    const size_t a2_bytes = ext2 > 0 ? (size_t)(ext2 - 1) * angle2_stride + sizeof(float) : 0;
    size_t o = 0;
    const size_t o_d1 = o; o += al((size_t)32 * std::max(ext1, 1));
    const size_t o_d2 = o; o += al((size_t)32 * std::max(ext2, 1));
    const size_t o_a1 = o; o += al(std::max(a1_bytes, (size_t)4));
    const size_t o_a2 = o; o += al(std::max(a2_bytes, (size_t)4));
    const size_t o_v2 = o; o += al((size_t)std::max(ext2, 1));
    const size_t o_mt = o; o += al(sizeof(int) * meta.size());
    const size_t o_pr = o; o += al(sizeof(int) * 2 * (size_t)pairs_stride * n_problems);
    const size_t o_np = o; o += al(sizeof(int) * (size_t)n_problems);
    int rc = m.grow((void**)&m.d_stage, &m.stage_cap, o);
    if (rc) return rc;
    unsigned char* s = m.d_stage;
    cudaStream_t st = m.stream;
    if (ext1 > 0) {
        B200_CUDA(cudaMemcpyAsync(s + o_d1, desc1 + (size_t)32 * lo1, (size_t)32 * ext1, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(s + o_a1, (const unsigned char*)angle1 + (size_t)lo1 * angle1_stride, a1_bytes, cudaMemcpyHostToDevice, st));
    }
    if (ext2 > 0) {
        B200_CUDA(cudaMemcpyAsync(s + o_d2, desc2 + (size_t)32 * lo2, (size_t)32 * ext2, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(s + o_a2, (const unsigned char*)angle2 + (size_t)lo2 * angle2_stride, a2_bytes, cudaMemcpyHostToDevice, st));
        if (valid2) B200_CUDA(cudaMemcpyAsync(s + o_v2, valid2 + lo2, ext2, cudaMemcpyHostToDevice, st));
    }
    B200_CUDA(cudaMemcpyAsync(s + o_mt, meta.data(), sizeof(int) * meta.size(), cudaMemcpyHostToDevice, st));
    const int* dm = (const int*)(s + o_mt);
    const Side S1{(const uint4*)(s + o_d1), s + o_a1, (long long)angle1_stride, dm, dm + n_problems};
    const Side S2{(const uint4*)(s + o_d2), s + o_a2, (long long)angle2_stride, dm + 2 * n_problems, dm + 3 * n_problems};
    rc = m.run(n_problems, S1, S2, valid2 ? s + o_v2 : nullptr, max_n1, max_n2, lowe_ratio, check_orientation, s + o_pr, pairs_stride,
               s + o_np);
    if (rc) return rc;
    m.last_h2d = (size_t)32 * (ext1 + ext2) + a1_bytes + a2_bytes + (valid2 ? ext2 : 0) + sizeof(int) * meta.size();
    m.last_d2h = sizeof(int) * ((size_t)n_problems + 2 * (size_t)pairs_stride * n_problems);
    B200_CUDA(cudaMemcpyAsync(n_pairs, s + o_np, sizeof(int) * n_problems, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(pairs, s + o_pr, sizeof(int) * 2 * (size_t)pairs_stride * n_problems, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
}

int b200_match_guided(b200_matcher_t h, int n_problems, b200_guided_problem_t* problems, int mode, unsigned thr, float lowe_ratio,
                      int check_orientation, int max_candidates) {
    B200_RANGE("b200:match:guided");
    using b200::match::GuidedDev;
    if (!h || n_problems < 0 || mode < B200_GUIDED_LANDMARKS || mode > B200_GUIDED_AREA || max_candidates < 0) return B200_ERR_INVALID;
    if (n_problems == 0) return B200_OK;
    if (!problems) return B200_ERR_INVALID;
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    const int cap = max_candidates ? max_candidates : 256;
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    // layout of the arena: [GuidedDev x n][inputs of every problem][outputs of every problem][scratch]; the first two parts are
    // mirrored in pinned host memory and go up in one copy, the output part comes back in one copy.
    struct Lay {
        size_t tx, ty, toct, tang, txr, tdesc, qdesc, qx, qy, qm, qlo, qhi, qxr, qang, qval, qrep, sig;  // inputs
        size_t occ, mout, nm;                                                                // outputs
        size_t cstart, citems, ccur, lists, llen, own;                                       // scratch
    };
    std::vector<Lay> lay(n_problems);
    size_t o = al(sizeof(GuidedDev) * (size_t)n_problems);
    int max_q = 0, max_train = 0;
    for (int p = 0; p < n_problems; ++p) {
        const b200_guided_problem_t& P = problems[p];
        const bool need_angle = (mode == B200_GUIDED_LAST_FRAME || mode == B200_GUIDED_AREA) && check_orientation;
        const bool need_reproj = mode == B200_GUIDED_FUSE && P.do_reprojection_matching;
        const bool need_xr = P.t_x_right && (mode <= B200_GUIDED_LAST_FRAME || need_reproj);
        if (P.n_train < 0 || P.n_queries < 0 || P.grid_cols <= 0 || P.grid_rows <= 0 || !(P.max_x > P.min_x) || !(P.max_y > P.min_y)
            || (long long)P.grid_cols * P.grid_rows > (1 << 20)) {
            b200::set_error("b200_match_guided: bad sizes / image bounds in problem %d", p);
            return B200_ERR_INVALID;
        }
        if ((P.n_train > 0 && (!P.t_x || !P.t_y || !P.t_octave || !P.t_desc || (need_angle && !P.t_angle)))
            || (P.n_queries > 0
                && (!P.q_desc || !P.q_x || !P.q_y || !P.q_margin || !P.q_min_level || !P.q_max_level || !P.match_out
                    || (need_angle && !P.q_angle) || (need_xr && !P.q_x_right)
                    || (need_reproj && (!P.q_reproj || !P.inv_level_sigma_sq || P.n_levels <= 0 || P.n_levels > 256))))) {
            b200::set_error("b200_match_guided: null buffer in problem %d", p);
            return B200_ERR_INVALID;
        }
        Lay& L = lay[p];
        const size_t nt = (size_t)std::max(P.n_train, 1), nq = (size_t)std::max(P.n_queries, 1);
        L.tx = o; o += al(4 * nt);
        L.ty = o; o += al(4 * nt);
        L.toct = o; o += al(nt);
        L.tang = o; o += al(4 * nt);
        L.txr = o; o += al(4 * nt);
        L.tdesc = o; o += al(32 * nt);
        L.qdesc = o; o += al(32 * nq);
        L.qx = o; o += al(4 * nq);
        L.qy = o; o += al(4 * nq);
        L.qm = o; o += al(4 * nq);
        L.qlo = o; o += al(nq);
        L.qhi = o; o += al(nq);
        L.qxr = o; o += al(4 * nq);
        L.qang = o; o += al(4 * nq);
        L.qval = o; o += al(nq);
        L.qrep = o; o += al(16 * nq);
        L.sig = o; o += al(4 * 256);
        L.occ = o; o += al(nt);  // in AND out: kept at the end of the problem's input block
        max_q = std::max(max_q, P.n_queries);
        max_train = std::max(max_train, P.n_train);
    }
    const size_t in_bytes = o;
    const size_t out_begin = o;
    for (int p = 0; p < n_problems; ++p) {
        Lay& L = lay[p];
        L.mout = o; o += al(4 * (size_t)std::max(problems[p].n_queries, 1));
        L.nm = o; o += al(4);
    }
    const size_t o_overflow = o;
    o += al(4);
    const size_t out_end = o;
    for (int p = 0; p < n_problems; ++p) {
        const b200_guided_problem_t& P = problems[p];
        Lay& L = lay[p];
        const size_t cells = (size_t)P.grid_cols * P.grid_rows;
        L.cstart = o; o += al(4 * (cells + 1));
        L.citems = o; o += al(4 * (size_t)std::max(P.n_train, 1));
        L.ccur = o; o += al(4 * cells);
        L.lists = o; o += al(8 * (size_t)cap * std::max(P.n_queries, 1));
        L.llen = o; o += al(4 * (size_t)std::max(P.n_queries, 1));
        L.own = o; o += al(4 * (size_t)std::max(P.n_train, 1));
    }
    const size_t rs_bytes = (size_t)std::max(max_train, 1) * 6 + 16;  // claim (int) + state (u16) per keypoint
    if (rs_bytes > 200 * 1024) {
        b200::set_error("b200_match_guided: %d keypoints per frame exceed the on-chip occupancy table", max_train);
        return B200_ERR_CAPACITY;
    }
    int rc;
    if ((rc = m.grow((void**)&m.d_guided, &m.d_guided_cap, o))) return rc;
    if ((rc = m.grow_pinned(&m.h_guided, &m.h_guided_cap, out_end))) return rc;
    unsigned char *hb = m.h_guided, *db = m.d_guided;
    GuidedDev* hg = reinterpret_cast<GuidedDev*>(hb);
    auto put = [&](size_t off, const void* src, size_t bytes) {
        if (src && bytes) std::memcpy(hb + off, src, bytes);
    };
    for (int p = 0; p < n_problems; ++p) {
        const b200_guided_problem_t& P = problems[p];
        const Lay& L = lay[p];
        const size_t nt = (size_t)P.n_train, nq = (size_t)P.n_queries;
        put(L.tx, P.t_x, 4 * nt);
        put(L.ty, P.t_y, 4 * nt);
        put(L.toct, P.t_octave, nt);
        put(L.tang, P.t_angle, 4 * nt);
        put(L.txr, P.t_x_right, 4 * nt);
        put(L.tdesc, P.t_desc, 32 * nt);
        put(L.qdesc, P.q_desc, 32 * nq);
        put(L.qx, P.q_x, 4 * nq);
        put(L.qy, P.q_y, 4 * nq);
        put(L.qm, P.q_margin, 4 * nq);
        put(L.qlo, P.q_min_level, nq);
        put(L.qhi, P.q_max_level, nq);
        put(L.qxr, P.q_x_right, 4 * nq);
        put(L.qang, P.q_angle, 4 * nq);
        if (P.q_valid || P.q_has_observation)  // bit 0: valid, bit 1: has_observation()
            for (int q = 0; q < nq; ++q)
                hb[L.qval + q] = (unsigned char)(((!P.q_valid || P.q_valid[q]) ? 1 : 0) | ((!P.q_has_observation || P.q_has_observation[q]) ? 2 : 0));
        const bool reproj = mode == B200_GUIDED_FUSE && P.do_reprojection_matching;
        if (reproj) {
            put(L.qrep, P.q_reproj, 16 * nq);
            std::memset(hb + L.sig, 0, 4 * 256);
            put(L.sig, P.inv_level_sigma_sq, 4 * (size_t)P.n_levels);
        }
        if (P.t_occupied && mode != B200_GUIDED_AREA) put(L.occ, P.t_occupied, nt);
        else std::memset(hb + L.occ, 0, nt);
        GuidedDev g{};
        g.n_train = P.n_train;
        g.n_queries = P.n_queries;
        g.grid_cols = P.grid_cols;
        g.grid_rows = P.grid_rows;
        g.cap = cap;
        g.min_x = P.min_x; g.max_x = P.max_x; g.min_y = P.min_y; g.max_y = P.max_y;
        g.t_x = (const float*)(db + L.tx);
        g.t_y = (const float*)(db + L.ty);
        g.t_angle = (const float*)(db + L.tang);
        g.t_x_right = P.t_x_right ? (const float*)(db + L.txr) : nullptr;
        g.t_octave = db + L.toct;
        g.t_desc = (const uint4*)(db + L.tdesc);
        g.q_desc = (const uint4*)(db + L.qdesc);
        g.q_x = (const float*)(db + L.qx);
        g.q_y = (const float*)(db + L.qy);
        g.q_margin = (const float*)(db + L.qm);
        g.q_x_right = (const float*)(db + L.qxr);
        g.q_angle = (const float*)(db + L.qang);
        g.q_min_level = (const signed char*)(db + L.qlo);
        g.q_max_level = (const signed char*)(db + L.qhi);
        g.q_valid = (P.q_valid || P.q_has_observation) ? db + L.qval : nullptr;
        g.q_reproj = (const double*)(db + L.qrep);
        g.inv_level_sigma_sq = (const float*)(db + L.sig);
        g.do_reproj = reproj ? 1 : 0;
        g.owner = (int*)(db + L.own);
        g.cell_start = (int*)(db + L.cstart);
        g.cell_items = (int*)(db + L.citems);
        g.cell_cursor = (int*)(db + L.ccur);
        g.lists = (uint2*)(db + L.lists);
        g.list_len = (int*)(db + L.llen);
        g.occupied = db + L.occ;
        g.match_out = (int*)(db + L.mout);
        g.n_matches = (int*)(db + L.nm);
        hg[p] = g;
    }
    cudaStream_t st = m.stream;
    B200_CUDA(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemsetAsync(db + o_overflow, 0, 4, st));
    const GuidedDev* dg = reinterpret_cast<const GuidedDev*>(db);
    b200::match::guided_grid_kernel<<<n_problems, 1024, 0, st>>>(dg);
    b200::match::guided_candidates_kernel<<<dim3(std::max(1, b200::ceil_div(max_q, 128)), n_problems), 128, 0, st>>>(dg, mode, check_orientation,
                                                                                                                     (int*)(db + o_overflow));
    if (rs_bytes > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(b200::match::guided_resolve_kernel<GuidedDev>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_bytes));
    b200::match::guided_resolve_kernel<GuidedDev><<<n_problems, 32, rs_bytes, st>>>(dg, mode, thr, lowe_ratio);
    B200_CUDA(cudaGetLastError());
    // outputs: occupancy lives in the input block (copied back per problem only when asked for), the rest is contiguous
    B200_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, st));
    size_t occ_bytes = 0;
    const bool writes_occupancy = mode == B200_GUIDED_LANDMARKS || mode == B200_GUIDED_LAST_FRAME || mode == B200_GUIDED_FUSE;
    for (int p = 0; p < n_problems; ++p)
        if (writes_occupancy && problems[p].t_occupied && problems[p].n_train > 0) {
            B200_CUDA(cudaMemcpyAsync(hb + lay[p].occ, db + lay[p].occ, (size_t)problems[p].n_train, cudaMemcpyDeviceToHost, st));
            occ_bytes += (size_t)problems[p].n_train;
        }
    B200_CUDA(cudaStreamSynchronize(st));
    m.last_h2d = in_bytes;
    m.last_d2h = out_end - out_begin + occ_bytes;
    const int overflow = *reinterpret_cast<const int*>(hb + o_overflow);
    if (overflow > 0) {
        b200::set_error("b200_match_guided: a search window returned %d keypoints, max_candidates is %d", overflow, cap);
        return B200_ERR_CAPACITY;
    }
    for (int p = 0; p < n_problems; ++p) {
        b200_guided_problem_t& P = problems[p];
        if (P.n_queries > 0) std::memcpy(P.match_out, hb + lay[p].mout, 4 * (size_t)P.n_queries);
        P.n_matches = *reinterpret_cast<const int*>(hb + lay[p].nm);
        if (writes_occupancy && P.t_occupied && P.n_train > 0) std::memcpy(P.t_occupied, hb + lay[p].occ, (size_t)P.n_train);
    }
    return B200_OK;
}

int b200_track_stage_ms(b200_matcher_t h, int stage, float* ms) {
    if (!h || !ms || stage < 0 || stage > 6 || !h->m.track_timed) return B200_ERR_INVALID;
    if (stage == 6) B200_CUDA(cudaEventElapsedTime(ms, h->m.ev_track[0], h->m.ev_track[6]));
    else B200_CUDA(cudaEventElapsedTime(ms, h->m.ev_track[stage], h->m.ev_track[stage + 1]));
    return B200_OK;
}

// The device-resident tracking chain (see include/b200vslam.h).  Arena of the matcher handle:
//   [TrackFrameDev x n][GuidedDev x n][caller inputs]  -- mirrored in pinned memory, one upload
//   [outputs]                                          -- one download
//   [stage-A products, guided scratch]
int b200_track_local_map(b200_orb_t orb, b200_matcher_t h, b200_lba_t opt, const b200_track_params_t* prm, int n_frames, b200_track_frame_t* frames) {
    B200_RANGE("b200:track:local_map");
    using b200::chain::TrackFrameDev;
    using b200::chain::TrackShared;
    using b200::match::GuidedDev;
    if (!orb || !h || !opt || !prm || n_frames < 0) return B200_ERR_INVALID;
    if (n_frames == 0) return B200_OK;
    if (!frames || !prm->scale_factors || !prm->inv_level_sigma_sq || prm->num_levels == 0 || prm->num_levels > 32 || prm->grid_cols <= 0
        || prm->grid_rows <= 0 || (long long)prm->grid_cols * prm->grid_rows > (1 << 20) || !(prm->img_bounds[1] > prm->img_bounds[0])
        || !(prm->img_bounds[3] > prm->img_bounds[2]) || (prm->cam.model != 0 && prm->cam.model != 1) || prm->max_candidates < 0
        || prm->num_trials_robust < 0 || prm->num_trials < 0 || prm->num_each_iter < 0) {
        b200::set_error("b200_track_local_map: invalid parameters");
        return B200_ERR_INVALID;
    }
    auto& m = h->m;
    const b200_keypoint_t* d_kps = nullptr;
    const unsigned char* d_descs = nullptr;
    const int* d_counts = nullptr;
    int stride = 0, batch = 0, device = 0;
    cudaStream_t st = nullptr;
    int rc = b200::chain::orb_results(orb, &d_kps, &d_descs, &d_counts, &stride, &batch, &st, &device);
    if (rc) return rc;
    if (device != m.device) {
        b200::set_error("b200_track_local_map: extractor and matcher live on different devices");
        return B200_ERR_INVALID;
    }
    B200_CUDA(cudaSetDevice(m.device));
    const int cap = prm->max_candidates ? prm->max_candidates : 256;
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    struct Lay {
        size_t xr, kl, pos, nrm, lo, hi, desc, skip, hobs;                       // inputs
        size_t obs, klo, kout, status, mout, nm;                                   // outputs
        size_t und, tx, ty, toct, occ, qx, qy, qm, qxr, qlo, qhi, qval;           // stage A products
        size_t cstart, citems, ccur, lists, llen, own;                            // guided scratch
    };
    std::vector<Lay> lay(n_frames);
    int max_lm = 0;
    size_t o = al(sizeof(TrackFrameDev) * (size_t)n_frames);
    const size_t o_gd = o;
    o += al(sizeof(GuidedDev) * (size_t)n_frames);
    for (int f = 0; f < n_frames; ++f) {
        const b200_track_frame_t& F = frames[f];
        if (F.frame < 0 || F.frame >= batch || !F.pose_cw || F.n_landmarks < 0 || F.n_keypoints_in < 0 || F.kp_cap < 0
            || ((F.kp_x_right || F.kp_landmark) && F.n_keypoints_in > stride) || !F.kp_landmark_out || !F.kp_outlier
            || (F.n_landmarks > 0 && (!F.lm_pos_w || !F.lm_mean_normal || !F.lm_min_valid_dist || !F.lm_max_valid_dist || !F.lm_desc || !F.lm_observable))) {
            b200::set_error("b200_track_local_map: frame %d: bad frame index, sizes or null buffers", f);
            return B200_ERR_INVALID;
        }
        Lay& L = lay[f];
        const size_t nk = (size_t)std::max(F.n_keypoints_in, 1), nl = (size_t)std::max(F.n_landmarks, 1);
        L.xr = o; o += al(4 * nk);
        L.kl = o; o += al(4 * nk);
        L.pos = o; o += al(24 * nl);
        L.nrm = o; o += al(24 * nl);
        L.lo = o; o += al(4 * nl);
        L.hi = o; o += al(4 * nl);
        L.desc = o; o += al(32 * nl);
        L.skip = o; o += al(nl);
        L.hobs = o; o += al(nl);
        max_lm = std::max(max_lm, F.n_landmarks);
    }
    const size_t in_bytes = o, out_begin = o;
    const size_t kc = (size_t)std::max(stride, 1);
    for (int f = 0; f < n_frames; ++f) {
        Lay& L = lay[f];
        const size_t nl = (size_t)std::max(frames[f].n_landmarks, 1);
        L.obs = o; o += al(nl);
        L.klo = o; o += al(4 * kc);
        L.kout = o; o += al(kc);
        L.status = o; o += al(16);
        L.mout = o; o += al(4 * nl);
        L.nm = o; o += al(4);
    }
    const size_t o_pose = o; o += al(8 * 16 * (size_t)n_frames);
    const size_t o_nvalid = o; o += al(4 * (size_t)n_frames);
    const size_t o_overflow = o; o += al(4);
    const size_t out_end = o;
    const size_t cells = (size_t)prm->grid_cols * prm->grid_rows;
    for (int f = 0; f < n_frames; ++f) {
        Lay& L = lay[f];
        const size_t nl = (size_t)std::max(frames[f].n_landmarks, 1);
        L.und = o; o += al(sizeof(b200_keypoint_t) * kc);
        L.tx = o; o += al(4 * kc);
        L.ty = o; o += al(4 * kc);
        L.toct = o; o += al(kc);
        L.occ = o; o += al(kc);
        L.qx = o; o += al(4 * nl);
        L.qy = o; o += al(4 * nl);
        L.qm = o; o += al(4 * nl);
        L.qxr = o; o += al(4 * nl);
        L.qlo = o; o += al(nl);
        L.qhi = o; o += al(nl);
        L.qval = o; o += al(nl);
        L.cstart = o; o += al(4 * (cells + 1));
        L.citems = o; o += al(4 * kc);
        L.ccur = o; o += al(4 * cells);
        L.lists = o; o += al(8 * (size_t)cap * nl);
        L.llen = o; o += al(4 * nl);
        L.own = o; o += al(4 * kc);
    }
    const size_t rs_bytes = kc * 6 + 16;
    if (rs_bytes > 200 * 1024) {
        b200::set_error("b200_track_local_map: %d keypoints per frame exceed the on-chip occupancy table", stride);
        return B200_ERR_CAPACITY;
    }
    if ((rc = m.grow((void**)&m.d_guided, &m.d_guided_cap, o))) return rc;
    if ((rc = m.grow_pinned(&m.h_guided, &m.h_guided_cap, out_end))) return rc;
    for (int i = 0; i < 7; ++i)
        if (!m.ev_track[i]) B200_CUDA(cudaEventCreate(&m.ev_track[i]));
    unsigned char *hb = m.h_guided, *db = m.d_guided;
    TrackFrameDev* hf = reinterpret_cast<TrackFrameDev*>(hb);
    GuidedDev* hg = reinterpret_cast<GuidedDev*>(hb + o_gd);
    auto put = [&](size_t off, const void* src, size_t bytes) {
        if (src && bytes) std::memcpy(hb + off, src, bytes);
    };
    TrackShared sh{};
    sh.model = prm->cam.model;
    sh.fx = prm->cam.fx; sh.fy = prm->cam.fy; sh.cx = prm->cam.cx; sh.cy = prm->cam.cy;
    sh.k1 = prm->cam.k1; sh.k2 = prm->cam.k2; sh.p1 = prm->cam.p1; sh.p2 = prm->cam.p2; sh.k3 = prm->cam.k3;
    sh.cols = prm->cam.cols; sh.rows = prm->cam.rows;
    sh.fxb = prm->focal_x_baseline;
    sh.min_x = prm->img_bounds[0]; sh.max_x = prm->img_bounds[1]; sh.min_y = prm->img_bounds[2]; sh.max_y = prm->img_bounds[3];
    sh.ray_cos_thr = prm->ray_cos_thr;
    sh.log_scale_factor = prm->log_scale_factor;
    sh.margin = prm->margin;
    sh.delta = prm->monocular ? std::sqrt(5.99146f) : std::sqrt(7.81473f);  // pose_optimizer_g2o.cc:73-88
    sh.num_levels = prm->num_levels;
    for (unsigned l = 0; l < 32; ++l) {
        sh.scale_factors[l] = prm->scale_factors[std::min(l, prm->num_levels - 1)];
        sh.inv_level_sigma_sq[l] = prm->inv_level_sigma_sq[std::min(l, prm->num_levels - 1)];
    }
    std::vector<const double*> poses(n_frames);
    for (int f = 0; f < n_frames; ++f) {
        const b200_track_frame_t& F = frames[f];
        const Lay& L = lay[f];
        const size_t nk = (size_t)F.n_keypoints_in, nl = (size_t)F.n_landmarks;
        put(L.xr, F.kp_x_right, 4 * nk);
        put(L.kl, F.kp_landmark, 4 * nk);
        put(L.pos, F.lm_pos_w, 24 * nl);
        put(L.nrm, F.lm_mean_normal, 24 * nl);
        put(L.lo, F.lm_min_valid_dist, 4 * nl);
        put(L.hi, F.lm_max_valid_dist, 4 * nl);
        put(L.desc, F.lm_desc, 32 * nl);
        put(L.skip, F.lm_skip, nl);
        put(L.hobs, F.lm_has_observation, nl);
        poses[f] = F.pose_cw;
        TrackFrameDev t{};
        t.kps = d_kps + (size_t)F.frame * stride;
        t.n_kp = d_counts + F.frame;
        t.kp_cap = stride;
        t.n_kp_in = F.n_keypoints_in;
        t.n_lm = F.n_landmarks;
        t.kp_x_right = F.kp_x_right ? (const float*)(db + L.xr) : nullptr;
        t.kp_landmark = F.kp_landmark ? (const int*)(db + L.kl) : nullptr;
        t.pos_w = (const double*)(db + L.pos);
        t.mean_normal = (const double*)(db + L.nrm);
        t.min_d = (const float*)(db + L.lo);
        t.max_d = (const float*)(db + L.hi);
        t.lm_skip = F.lm_skip ? db + L.skip : nullptr;
        t.lm_has_obs = F.lm_has_observation ? db + L.hobs : nullptr;
        const double* P = F.pose_cw;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) t.Rt[3 * r + c] = P[4 * r + c];
            t.Rt[9 + r] = P[4 * r + 3];
        }
        for (int r = 0; r < 3; ++r) t.twc[r] = -(t.Rt[r] * t.Rt[9] + t.Rt[3 + r] * t.Rt[10] + t.Rt[6 + r] * t.Rt[11]);
        t.undist = (b200_keypoint_t*)(db + L.und);
        t.t_x = (float*)(db + L.tx);
        t.t_y = (float*)(db + L.ty);
        t.t_octave = db + L.toct;
        t.occupied = db + L.occ;
        t.observable = db + L.obs;
        t.q_x = (float*)(db + L.qx);
        t.q_y = (float*)(db + L.qy);
        t.q_margin = (float*)(db + L.qm);
        t.q_xr = (float*)(db + L.qxr);
        t.q_lo = (signed char*)(db + L.qlo);
        t.q_hi = (signed char*)(db + L.qhi);
        t.q_valid = db + L.qval;
        t.match_out = (const int*)(db + L.mout);
        t.kp_landmark_out = (int*)(db + L.klo);
        t.kp_outlier = db + L.kout;
        t.status = (int*)(db + L.status);
        hf[f] = t;
        GuidedDev g{};
        g.n_train = 0;  // set on the device from the extractor's counter
        g.n_queries = F.n_landmarks;
        g.grid_cols = prm->grid_cols;
        g.grid_rows = prm->grid_rows;
        g.cap = cap;
        g.min_x = sh.min_x; g.max_x = sh.max_x; g.min_y = sh.min_y; g.max_y = sh.max_y;
        g.t_x = t.t_x;
        g.t_y = t.t_y;
        g.t_angle = nullptr;
        g.t_x_right = t.kp_x_right;
        g.t_octave = t.t_octave;
        g.t_desc = reinterpret_cast<const uint4*>(d_descs + (size_t)F.frame * stride * 32);
        g.q_desc = (const uint4*)(db + L.desc);
        g.q_x = t.q_x;
        g.q_y = t.q_y;
        g.q_margin = t.q_margin;
        g.q_x_right = t.q_xr;
        g.q_angle = nullptr;
        g.q_min_level = t.q_lo;
        g.q_max_level = t.q_hi;
        g.q_valid = t.q_valid;
        g.q_reproj = nullptr;
        g.inv_level_sigma_sq = nullptr;
        g.do_reproj = 0;
        g.owner = (int*)(db + L.own);
        g.cell_start = (int*)(db + L.cstart);
        g.cell_items = (int*)(db + L.citems);
        g.cell_cursor = (int*)(db + L.ccur);
        g.lists = (uint2*)(db + L.lists);
        g.list_len = (int*)(db + L.llen);
        g.occupied = t.occupied;
        g.match_out = (int*)(db + L.mout);
        g.n_matches = (int*)(db + L.nm);
        hg[f] = g;
    }
    if ((rc = m.join())) return rc;
    B200_CUDA(cudaEventRecord(m.ev_track[0], st));
    B200_CUDA(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemsetAsync(db + out_begin, 0, out_end - out_begin, st));
    const TrackFrameDev* df = reinterpret_cast<const TrackFrameDev*>(db);
    GuidedDev* dg = reinterpret_cast<GuidedDev*>(db + o_gd);
    if ((rc = b200::chain::track_stage_a(st, sh, df, n_frames, stride, max_lm))) return rc;
    b200::match::track_set_counts_kernel<<<b200::ceil_div(n_frames, 128), 128, 0, st>>>(dg, df, n_frames);
    B200_CUDA(cudaEventRecord(m.ev_track[1], st));
    b200::match::guided_grid_kernel<<<n_frames, 1024, 0, st>>>(dg);
    B200_CUDA(cudaEventRecord(m.ev_track[2], st));
    b200::match::guided_candidates_kernel<<<dim3(std::max(1, b200::ceil_div(max_lm, 128)), n_frames), 128, 0, st>>>(dg, B200_GUIDED_LANDMARKS, 0,
                                                                                                                     (int*)(db + o_overflow));
    B200_CUDA(cudaEventRecord(m.ev_track[3], st));
    if (rs_bytes > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(b200::match::guided_resolve_kernel<GuidedDev>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_bytes));
    b200::match::guided_resolve_kernel<GuidedDev><<<n_frames, 32, rs_bytes, st>>>(dg, B200_GUIDED_LANDMARKS, prm->hamming_thr, prm->lowe_ratio);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaEventRecord(m.ev_track[4], st));
    if ((rc = b200::chain::track_stage_c(opt, st, sh, df, hf, poses.data(), n_frames, stride, prm->num_trials_robust, prm->num_trials, prm->num_each_iter,
                                         (double*)(db + o_pose), (unsigned*)(db + o_nvalid), m.ev_track[5])))
        return rc;
    B200_CUDA(cudaEventRecord(m.ev_track[6], st));
    B200_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    m.track_timed = true;
    m.last_h2d = in_bytes;
    m.last_d2h = out_end - out_begin;
    const int overflow = *reinterpret_cast<const int*>(hb + o_overflow);
    if (overflow > 0) {
        b200::set_error("b200_track_local_map: a search window returned %d keypoints, max_candidates is %d", overflow, cap);
        return B200_ERR_CAPACITY;
    }
    for (int f = 0; f < n_frames; ++f) {
        b200_track_frame_t& F = frames[f];
        const Lay& L = lay[f];
        const int* status = reinterpret_cast<const int*>(hb + L.status);
        const int nk = status[0];
        if (status[1]) {
            b200::set_error("b200_track_local_map: frame %d has %d keypoints, the caller's per-keypoint arrays have %d", f, nk, F.n_keypoints_in);
            return B200_ERR_INVALID;
        }
        if (nk > F.kp_cap) {
            b200::set_error("b200_track_local_map: frame %d has %d keypoints, kp_cap is %d", f, nk, F.kp_cap);
            return B200_ERR_CAPACITY;
        }
        F.n_keypoints = nk;
        if (F.n_landmarks > 0) std::memcpy(F.lm_observable, hb + L.obs, (size_t)F.n_landmarks);
        std::memcpy(F.kp_landmark_out, hb + L.klo, 4 * (size_t)nk);
        std::memcpy(F.kp_outlier, hb + L.kout, (size_t)nk);
        F.n_matches = *reinterpret_cast<const int*>(hb + L.nm);
        F.n_valid = reinterpret_cast<const unsigned*>(hb + o_nvalid)[f];
        if (status[2] < 5) std::memcpy(F.pose_cw_out, F.pose_cw, sizeof(double) * 16);  // :116-118: returns before touching the pose
        else std::memcpy(F.pose_cw_out, hb + o_pose + sizeof(double) * 16 * (size_t)f, sizeof(double) * 16);
    }
    return B200_OK;
}

int b200_match_pairs(b200_matcher_t h, int n_problems, b200_pairs_problem_t* problems, int variant, float lowe_ratio, int check_orientation,
                     int max_candidates) {
    B200_RANGE("b200:match:pairs");
    using b200::match::PairsDev;
    if (!h || n_problems < 0 || (variant != B200_PAIRS_BOW && variant != B200_PAIRS_TRIANGULATION) || max_candidates < 0) return B200_ERR_INVALID;
    if (n_problems == 0) return B200_OK;
    if (!problems) return B200_ERR_INVALID;
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    const int cap = max_candidates ? max_candidates : 64;
    const bool tri = variant == B200_PAIRS_TRIANGULATION;
    // distances that can still matter: the match itself needs <= 50; a second-best above T can no longer fail the ratio test
    // (lowe * (T + 1) >= 50 >= best, bow_tree.cc:232-239).  Triangulation never looks above 50 (robust.cc:83-85).
    unsigned list_thr = b200::match::kThrLow;
    if (!tri)
        while (list_thr < 255u && lowe_ratio * (float)(list_thr + 1) < (float)b200::match::kThrLow) ++list_thr;
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    struct Item {
        size_t off;
        const void* src;
        size_t bytes;
    };
    std::vector<Item> items;
    size_t o = al(sizeof(PairsDev) * (size_t)n_problems);
    auto add = [&](const void* src, size_t bytes) {
        const size_t at = o;
        items.push_back({at, src, src ? bytes : 0});
        o += al(std::max(bytes, (size_t)1));
        return at;
    };
    struct Lay {
        size_t d1, a1, v1, nd1, b1, s1, st1, d2, a2, v2, nd2, b2, st2, mout, nm, lists, llen;
    };
    std::vector<Lay> lay(n_problems);
    int max_n1 = 0, max_n2 = 0;
    for (int p = 0; p < n_problems; ++p) {
        const b200_pairs_problem_t& P = problems[p];
        if (P.n1 < 0 || P.n2 < 0 || (P.n1 > 0 && (!P.desc1 || !P.match_out)) || (P.n2 > 0 && !P.desc2) || ((P.node1 == nullptr) != (P.node2 == nullptr))
            || (check_orientation && ((P.n1 > 0 && !P.angle1) || (P.n2 > 0 && !P.angle2)))
            || (tri && ((P.n1 > 0 && (!P.bearing1 || !P.scale1)) || (P.n2 > 0 && !P.bearing2)))) {
            b200::set_error("b200_match_pairs: bad sizes or null buffer in problem %d", p);
            return B200_ERR_INVALID;
        }
        Lay& L = lay[p];
        const size_t n1 = (size_t)P.n1, n2 = (size_t)P.n2;
        L.d1 = add(P.desc1, 32 * n1);
        L.a1 = add(check_orientation ? P.angle1 : nullptr, 4 * n1);
        L.v1 = add(P.valid1, n1);
        L.nd1 = add(P.node1, 4 * n1);
        L.b1 = add(tri ? P.bearing1 : nullptr, 24 * n1);
        L.s1 = add(tri ? P.scale1 : nullptr, 4 * n1);
        L.st1 = add(tri ? P.stereo1 : nullptr, n1);
        L.d2 = add(P.desc2, 32 * n2);
        L.a2 = add(check_orientation ? P.angle2 : nullptr, 4 * n2);
        L.v2 = add(P.valid2, n2);
        L.nd2 = add(P.node2, 4 * n2);
        L.b2 = add(tri ? P.bearing2 : nullptr, 24 * n2);
        L.st2 = add(tri ? P.stereo2 : nullptr, n2);
        max_n1 = std::max(max_n1, P.n1);
        max_n2 = std::max(max_n2, P.n2);
    }
    const size_t in_bytes = o, out_begin = o;
    for (int p = 0; p < n_problems; ++p) {
        lay[p].mout = o; o += al(4 * (size_t)std::max(problems[p].n1, 1));
        lay[p].nm = o; o += al(4);
    }
    const size_t o_overflow = o;
    o += al(4);
    const size_t out_end = o;
    for (int p = 0; p < n_problems; ++p) {
        lay[p].lists = o; o += al(8 * (size_t)cap * std::max(problems[p].n1, 1));
        lay[p].llen = o; o += al(4 * (size_t)std::max(problems[p].n1, 1));
    }
    const size_t rs_bytes = (size_t)std::max(max_n2, 1) * 6 + 16;
    if (rs_bytes > 200 * 1024) {
        b200::set_error("b200_match_pairs: %d keypoints per frame exceed the on-chip occupancy table", max_n2);
        return B200_ERR_CAPACITY;
    }
    int rc;
    if ((rc = m.grow((void**)&m.d_guided, &m.d_guided_cap, o))) return rc;
    if ((rc = m.grow_pinned(&m.h_guided, &m.h_guided_cap, out_end))) return rc;
    unsigned char *hb = m.h_guided, *db = m.d_guided;
    for (const Item& it : items)
        if (it.bytes) std::memcpy(hb + it.off, it.src, it.bytes);
    PairsDev* hg = reinterpret_cast<PairsDev*>(hb);
    for (int p = 0; p < n_problems; ++p) {
        const b200_pairs_problem_t& P = problems[p];
        const Lay& L = lay[p];
        PairsDev g{};
        g.n_queries = P.n1;
        g.n_train = P.n2;
        g.cap = cap;
        g.desc1 = (const uint4*)(db + L.d1);
        g.desc2 = (const uint4*)(db + L.d2);
        g.angle1 = (const float*)(db + L.a1);
        g.angle2 = (const float*)(db + L.a2);
        g.valid1 = P.valid1 ? db + L.v1 : nullptr;
        g.valid2 = P.valid2 ? db + L.v2 : nullptr;
        g.stereo1 = (tri && P.stereo1) ? db + L.st1 : nullptr;
        g.stereo2 = (tri && P.stereo2) ? db + L.st2 : nullptr;
        g.node1 = P.node1 ? (const int*)(db + L.nd1) : nullptr;
        g.node2 = P.node2 ? (const int*)(db + L.nd2) : nullptr;
        g.bearing1 = (const double*)(db + L.b1);
        g.bearing2 = (const double*)(db + L.b2);
        g.scale1 = (const float*)(db + L.s1);
        for (int k = 0; k < 9; ++k) g.E[k] = P.E_12[k];
        for (int k = 0; k < 3; ++k) g.epi[k] = P.epiplane_in_keyfrm_2[k];
        g.valid_epiplane = P.valid_epiplane;
        g.residual_rad_thr = P.residual_rad_thr;
        g.lists = (uint2*)(db + L.lists);
        g.list_len = (int*)(db + L.llen);
        g.occupied = nullptr;
        g.match_out = (int*)(db + L.mout);
        g.n_matches = (int*)(db + L.nm);
        g.owner = nullptr;
        hg[p] = g;
    }
    cudaStream_t st = m.stream;
    B200_CUDA(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemsetAsync(db + o_overflow, 0, 4, st));
    const PairsDev* dg = reinterpret_cast<const PairsDev*>(db);
    b200::match::pairs_candidates_kernel<<<dim3(std::max(1, b200::ceil_div(max_n1, b200::match::kPairRows)), n_problems), b200::match::kPairRows, 0,
                                           st>>>(dg, variant, list_thr, check_orientation, (int*)(db + o_overflow));
    if (rs_bytes > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(b200::match::guided_resolve_kernel<PairsDev>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_bytes));
    b200::match::guided_resolve_kernel<PairsDev><<<n_problems, 32, rs_bytes, st>>>(dg, tri ? 6 : 5, (unsigned)b200::match::kThrLow, lowe_ratio);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    m.last_h2d = in_bytes;
    m.last_d2h = out_end - out_begin;
    const int overflow = *reinterpret_cast<const int*>(hb + o_overflow);
    if (overflow > 0) {
        b200::set_error("b200_match_pairs: a row kept %d gated candidates, max_candidates is %d", overflow, cap);
        return B200_ERR_CAPACITY;
    }
    for (int p = 0; p < n_problems; ++p) {
        b200_pairs_problem_t& P = problems[p];
        if (P.n1 > 0) std::memcpy(P.match_out, hb + lay[p].mout, 4 * (size_t)P.n1);
        P.n_matches = *reinterpret_cast<const int*>(hb + lay[p].nm);
    }
    return B200_OK;
}

int b200_stereo_compute(b200_matcher_t h, b200_orb_t left, int frame_left, b200_orb_t right, int frame_right, const b200_keypoint_t* keypts_left,
                        const uint8_t* descs_left, int n_left, const b200_keypoint_t* keypts_right, const uint8_t* descs_right, int n_right,
                        float focal_x_baseline, float true_baseline, float* stereo_x_right, float* depths, int32_t* n_matched) {
    B200_RANGE("b200:match:stereo");
    using namespace b200::match;
    if (!h || !left || !right || n_left < 0 || n_right < 0 || (n_left > 0 && (!keypts_left || !descs_left || !stereo_x_right || !depths))
        || (n_right > 0 && (!keypts_right || !descs_right)) || !(true_baseline > 0.f)) {
        b200::set_error("b200_stereo_compute: null argument, negative count or non-positive baseline");
        return B200_ERR_INVALID;
    }
    if (n_matched) *n_matched = 0;
    if (n_left == 0) return B200_OK;
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    StereoDev g{};
    int rc;
    for (int l = 0; l < kStereoMaxLevels; ++l) {
        int wl = 0, hl = 0, wr = 0, hr = 0;
        float sf = 0.f;
        if (b200_orb_level_info(left, l, &wl, &hl, nullptr, &sf) != B200_OK) break;
        if (b200_orb_level_info(right, l, &wr, &hr, nullptr, nullptr) != B200_OK || wl != wr || hl != hr) {
            b200::set_error("b200_stereo_compute: the two extractors hold different pyramids at level %d", l);
            return B200_ERR_INVALID;
        }
        StereoLevel& L = g.lv[l];
        size_t pl = 0, pr = 0;
        if ((rc = b200_orb_pyramid_level_view(left, frame_left, l, &L.left, &pl, &L.w, &L.h))) return rc;
        if ((rc = b200_orb_pyramid_level_view(right, frame_right, l, &L.right, &pr, nullptr, nullptr))) return rc;
        L.pitch_l = pl;
        L.pitch_r = pr;
        L.sf = sf;
        g.n_levels = l + 1;
    }
    if (g.n_levels == 0) {
        b200::set_error("b200_stereo_compute: the extractors hold no pyramid (run extract first)");
        return B200_ERR_INVALID;
    }
    // inv_scale_factors_: the float recurrence of orb_params.cc:37-48, not 1 / sf
    {
        const float inv1 = 1.0f / g.lv[g.n_levels > 1 ? 1 : 0].sf;
        g.lv[0].inv_sf = 1.0f;
        for (int l = 1; l < g.n_levels; ++l) g.lv[l].inv_sf = inv1 * g.lv[l - 1].inv_sf;
    }
    for (int i = 0; i < n_left; ++i)
        if (keypts_left[i].octave < 0 || keypts_left[i].octave >= g.n_levels) {
            b200::set_error("b200_stereo_compute: left keypoint %d has octave %d", i, keypts_left[i].octave);
            return B200_ERR_INVALID;
        }
    if ((rc = b200_orb_sync(left)) || (rc = b200_orb_sync(right))) return rc;
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    size_t o = al(sizeof(StereoDev));
    const size_t o_kl = o; o += al(sizeof(b200_keypoint_t) * (size_t)n_left);
    const size_t o_kr = o; o += al(sizeof(b200_keypoint_t) * (size_t)std::max(n_right, 1));
    const size_t o_dl = o; o += al((size_t)32 * n_left);
    const size_t o_dr = o; o += al((size_t)32 * std::max(n_right, 1));
    const size_t in_bytes = o, out_begin = o;
    const size_t o_x = o; o += al(4 * (size_t)n_left);
    const size_t o_dep = o; o += al(4 * (size_t)n_left);
    const size_t o_nk = o; o += al(4);
    const size_t out_end = o;
    const size_t o_best = o; o += al(4 * (size_t)n_left);
    const size_t o_corr = o; o += al(4 * (size_t)n_left);
    if ((rc = m.grow((void**)&m.d_guided, &m.d_guided_cap, o))) return rc;
    if ((rc = m.grow_pinned(&m.h_guided, &m.h_guided_cap, out_end))) return rc;
    unsigned char *hb = m.h_guided, *db = m.d_guided;
    g.n_left = n_left;
    g.n_right = n_right;
    g.kl = (const b200_keypoint_t*)(db + o_kl);
    g.kr = (const b200_keypoint_t*)(db + o_kr);
    g.dl = (const uint4*)(db + o_dl);
    g.dr = (const uint4*)(db + o_dr);
    g.fxb = focal_x_baseline;
    g.max_disp = focal_x_baseline / true_baseline;  // stereo.cc:18
    g.best_right = (int*)(db + o_best);
    g.x_right = (float*)(db + o_x);
    g.depth = (float*)(db + o_dep);
    g.corr = (int*)(db + o_corr);
    g.n_kept = (int*)(db + o_nk);
    std::memcpy(hb, &g, sizeof(g));
    std::memcpy(hb + o_kl, keypts_left, sizeof(b200_keypoint_t) * (size_t)n_left);
    std::memcpy(hb + o_dl, descs_left, (size_t)32 * n_left);
    if (n_right > 0) {
        std::memcpy(hb + o_kr, keypts_right, sizeof(b200_keypoint_t) * (size_t)n_right);
        std::memcpy(hb + o_dr, descs_right, (size_t)32 * n_right);
    }
    cudaStream_t st = m.stream;
    B200_CUDA(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
    const StereoDev* dg = reinterpret_cast<const StereoDev*>(db);
    stereo_match_kernel<<<b200::ceil_div(n_left, kStereoRows), kStereoRows, 0, st>>>(dg);
    stereo_subpixel_kernel<<<b200::ceil_div(n_left, 4), 128, 0, st>>>(dg);
    stereo_median_kernel<<<1, 1024, 0, st>>>(dg);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    m.last_h2d = in_bytes;
    m.last_d2h = out_end - out_begin;
    std::memcpy(stereo_x_right, hb + o_x, 4 * (size_t)n_left);
    std::memcpy(depths, hb + o_dep, 4 * (size_t)n_left);
    if (n_matched) *n_matched = *reinterpret_cast<const int*>(hb + o_nk);
    return B200_OK;
}

int b200_landmark_descriptors(b200_matcher_t h, int n_landmarks, const uint8_t* descs, const int32_t* offsets, int32_t* best_idx, uint8_t* desc_out) {
    if (!h || n_landmarks < 0) return B200_ERR_INVALID;
    if (n_landmarks == 0) return B200_OK;
    if (!offsets || !best_idx || offsets[0] != 0) {
        b200::set_error("b200_landmark_descriptors: offsets must start at 0 and best_idx must be given");
        return B200_ERR_INVALID;
    }
    for (int l = 0; l < n_landmarks; ++l)
        if (offsets[l + 1] < offsets[l]) {
            b200::set_error("b200_landmark_descriptors: offsets are not ascending at landmark %d", l);
            return B200_ERR_INVALID;
        }
    const size_t total = (size_t)offsets[n_landmarks];
    if (total > 0 && !descs) return B200_ERR_INVALID;
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    size_t o = 0;
    const size_t o_desc = o; o += al(32 * std::max(total, (size_t)1));
    const size_t o_off = o; o += al(4 * ((size_t)n_landmarks + 1));
    const size_t in_bytes = o, out_begin = o;
    const size_t o_best = o; o += al(4 * (size_t)n_landmarks);
    const size_t o_out = o; o += al(32 * (size_t)n_landmarks);
    const size_t out_end = o;
    int rc;
    if ((rc = m.grow((void**)&m.d_guided, &m.d_guided_cap, o))) return rc;
    if ((rc = m.grow_pinned(&m.h_guided, &m.h_guided_cap, out_end))) return rc;
    unsigned char *hb = m.h_guided, *db = m.d_guided;
    if (total) std::memcpy(hb + o_desc, descs, 32 * total);
    std::memcpy(hb + o_off, offsets, 4 * ((size_t)n_landmarks + 1));
    cudaStream_t st = m.stream;
    B200_CUDA(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
    b200::match::landmark_descriptor_kernel<<<b200::ceil_div(n_landmarks, 4), 128, 0, st>>>((const uint4*)(db + o_desc), (const int*)(db + o_off), n_landmarks,
                                                                                         (int*)(db + o_best), desc_out ? (uint4*)(db + o_out) : nullptr);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    m.last_h2d = in_bytes;
    m.last_d2h = out_end - out_begin;
    std::memcpy(best_idx, hb + o_best, 4 * (size_t)n_landmarks);
    for (int l = 0; l < n_landmarks; ++l)
        if (best_idx[l] == -2) {
            b200::set_error("b200_landmark_descriptors: landmark %d has %d observations, at most %d are supported", l, offsets[l + 1] - offsets[l],
                            b200::match::kLmMaxObs);
            return B200_ERR_CAPACITY;
        }
    if (desc_out) {
        std::memcpy(desc_out, hb + o_out, 32 * (size_t)n_landmarks);
        for (int l = 0; l < n_landmarks; ++l)
            if (best_idx[l] < 0) std::memset(desc_out + 32 * (size_t)l, 0, 32);
    }
    return B200_OK;
}

int b200_landmark_geometry(b200_matcher_t h, int n_landmarks, const double* pos_w, const int32_t* offsets, const double* cam_centers,
                           const double* ref_center, const float* ref_scale_factor, float inv_scale_factor_last, double* mean_normal,
                           float* max_valid_dist, float* min_valid_dist) {
    if (!h || n_landmarks < 0) return B200_ERR_INVALID;
    if (n_landmarks == 0) return B200_OK;
    if (!pos_w || !offsets || !ref_center || !ref_scale_factor || !mean_normal || !max_valid_dist || !min_valid_dist || offsets[0] != 0) return B200_ERR_INVALID;
    for (int l = 0; l < n_landmarks; ++l)
        if (offsets[l + 1] < offsets[l]) return B200_ERR_INVALID;
    const size_t N = (size_t)n_landmarks, total = (size_t)offsets[n_landmarks];
    if (total > 0 && !cam_centers) return B200_ERR_INVALID;
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    auto al = [](size_t v) { return b200::round_up(v, (size_t)256); };
    size_t o = 0;
    const size_t o_p = o; o += al(24 * N);
    const size_t o_off = o; o += al(4 * (N + 1));
    const size_t o_c = o; o += al(24 * std::max(total, (size_t)1));
    const size_t o_r = o; o += al(24 * N);
    const size_t o_s = o; o += al(4 * N);
    const size_t in_bytes = o, out_begin = o;
    const size_t o_mn = o; o += al(24 * N);
    const size_t o_mx = o; o += al(4 * N);
    const size_t o_mi = o; o += al(4 * N);
    const size_t out_end = o;
    int rc;
    if ((rc = m.grow((void**)&m.d_guided, &m.d_guided_cap, o))) return rc;
    if ((rc = m.grow_pinned(&m.h_guided, &m.h_guided_cap, out_end))) return rc;
    unsigned char *hb = m.h_guided, *db = m.d_guided;
    std::memcpy(hb + o_p, pos_w, 24 * N);
    std::memcpy(hb + o_off, offsets, 4 * (N + 1));
    if (total) std::memcpy(hb + o_c, cam_centers, 24 * total);
    std::memcpy(hb + o_r, ref_center, 24 * N);
    std::memcpy(hb + o_s, ref_scale_factor, 4 * N);
    cudaStream_t st = m.stream;
    B200_CUDA(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
    b200::match::landmark_geometry_kernel<<<b200::ceil_div(n_landmarks, 128), 128, 0, st>>>(n_landmarks, (const double*)(db + o_p), (const int*)(db + o_off),
                                                                                         (const double*)(db + o_c), (const double*)(db + o_r),
                                                                                         (const float*)(db + o_s), inv_scale_factor_last,
                                                                                         (double*)(db + o_mn), (float*)(db + o_mx), (float*)(db + o_mi));
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    std::memcpy(mean_normal, hb + o_mn, 24 * N);
    std::memcpy(max_valid_dist, hb + o_mx, 4 * N);
    std::memcpy(min_valid_dist, hb + o_mi, 4 * N);
    return B200_OK;
}

int b200_match_cross_check(const int32_t* idx2_in_1, int n1, const int32_t* idx1_in_2, int n2, int32_t* mutual_out, int32_t* n_mutual) {
    if (n1 < 0 || n2 < 0 || (n1 > 0 && (!idx2_in_1 || !mutual_out)) || (n2 > 0 && !idx1_in_2)) return B200_ERR_INVALID;
    int n = 0;
    for (int i = 0; i < n1; ++i) {
        const int j = idx2_in_1[i];
        const bool keep = 0 <= j && j < n2 && idx1_in_2[j] == i;
        mutual_out[i] = keep ? j : -1;
        n += keep;
    }
    if (n_mutual) *n_mutual = n;
    return B200_OK;
}

int b200_hamming_matrix(b200_matcher_t h, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, uint16_t* dist) {
    if (!h || n1 < 0 || n2 < 0) return B200_ERR_INVALID;
    if (n1 == 0 || n2 == 0) return B200_OK;
    if (!desc1 || !desc2 || !dist) return B200_ERR_INVALID;
    auto& m = h->m;
    B200_CUDA(cudaSetDevice(m.device));
    const size_t o_d2 = b200::round_up((size_t)32 * n1, (size_t)256), o_out = o_d2 + b200::round_up((size_t)32 * n2, (size_t)256);
    int rc = m.grow((void**)&m.d_stage, &m.stage_cap, o_out + sizeof(uint16_t) * (size_t)n1 * n2);
    if (rc) return rc;
    unsigned char* s = m.d_stage;
    B200_CUDA(cudaMemcpyAsync(s, desc1, (size_t)32 * n1, cudaMemcpyHostToDevice, m.stream));
    B200_CUDA(cudaMemcpyAsync(s + o_d2, desc2, (size_t)32 * n2, cudaMemcpyHostToDevice, m.stream));
    b200::match::hamming_matrix_kernel<<<dim3(b200::ceil_div(n2, 64), b200::ceil_div(n1, 256)), 256, 0, m.stream>>>(
        (const uint4*)s, n1, (const uint4*)(s + o_d2), n2, (unsigned short*)(s + o_out));
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(dist, s + o_out, sizeof(uint16_t) * (size_t)n1 * n2, cudaMemcpyDeviceToHost, m.stream));
    B200_CUDA(cudaStreamSynchronize(m.stream));
    return B200_OK;
}

}  // extern "C"
