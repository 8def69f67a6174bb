// blend_tc.cuh -- K2 (+K3): pose / shape blend as an FP16 GEMM on the 5th-gen tensor cores (tcgen05), with the
// LBS skinning optionally fused into its epilogue.
//
//   v_posed[f][c] = v_template[c] + sum_k feat[f][k] * dirs[k][c],   c in [0, 3V)
//   feat[f] = [ (R_1..R_{nj-1} - I).flatten() (9(nj-1)) | shape_hi | shape_lo | shape_hi | 0-pad ]
//   dirs    = [ posedirs ; S_hi ; S_hi ; S_lo ] * dir_scale
// The shape blend carries decimetre-scale offsets, so its rows are split into FP16 hi/lo parts on both
// sides (x = x_hi + x_lo, x_hi*S_hi + x_lo*S_hi + x_hi*S_lo; only the lo*lo term is dropped): that keeps the
// blend within ~1e-6 m of FP32.  FP16 has TF32's 10 mantissa bits but K = 16 per MMA instruction instead of
// 8, and the kernel is bound by MMA issue (~80 + 1.5 N cycles per instruction), so FP16 halves its time;
// dirs are scaled by a power of two into FP16's normal range and the epilogue multiplies by the inverse.
//
// smplx does this as `pose_feature @ posedirs` + `blend_shapes` [smplx-from-memory]; the reference
// reaches it through the final forward (/root/reference/keypoints2body/core/fitters/world_space.py:258-278).
//
// Orientation: D[M = 128 output coordinates][N = FR frames].  The M operand is a block of dirs^T
// (pre-tiled on the host into the exact shared-memory image of each (tile, 64-deep K block), streamed
// with ONE 1-D TMA bulk copy per block -- cp.async.bulk + mbarrier complete_tx -- into a ring); the N
// operand is the pass's frames' features, resident in shared memory.  Accumulators live in TENSOR MEMORY
// (4 x FR columns).  Two tiles are processed as a PAIR with their MMAs interleaved, and pairs alternate
// between accumulators {0,1} and {2,3} so the epilogue of one pair overlaps the MMAs of the next.  One thread
// issues tcgen05.mma.kind::f16 (M128 N<FR> K16).  Eight epilogue warps (lane quarter x tile of the pair)
// read TMEM with tcgen05.ld: lane = output coordinate, register i = frame.  Unfused (SMPL-H / SMPL-X,
// FR = 128): every store instruction writes 32 consecutive floats of one frame row of v_posed.  Fused
// (SMPL, FR = 64): see the epilogue -- tiles carry vertices, each lane skins its own coordinate.
// Warp roles synchronise only through mbarriers.
//
// Operand layout: K-major SWIZZLE_128B canonical UMMA layout (cute mma_traits_sm100: Swizzle<3,4,3>).
// A K block of 64 FP16 is one 128-byte row per operand row; rows are 128 B apart, 8-row groups
// 1024 B apart (SBO), and the 16-byte chunk index inside a row is XORed with (row % 8):
//   byte address(row r, k) = (k / 64) * rows*128 + r * 128 + ((((k % 64) / 8) ^ (r % 8)) * 16) + (k % 8) * 2
#pragma once

#ifdef K2B_DIAG
#define K2B_TC_DBG(p) ((p).debug)
#else
#define K2B_TC_DBG(p) 0     // work-skipping diagnostics do not exist in the shipped library
#endif

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace k2b {

constexpr int kTcM = 128;          // frames per pass (the MMA's N) of the plain blend (SMPL-H, SMPL-X; features <= 144 KB)
constexpr int kTcMFused = 64;      // frames per pass when the skinning rows share shared memory with the features (SMPL)
constexpr int kTcN = 128;          // output columns per tile (the MMA's M)
constexpr int kTcBK = 64;          // K block streamed per TMA copy = one 128-byte swizzle atom of FP16 (4 MMAs of K = 16)
constexpr int kTcKpadWide = 256;   // SMPL: 207 pose features + 3 x 10 split shape rows, padded to 8 x 32
constexpr int kTcKpadMax = 576;    // SMPL-X: 486 pose features + 3 x 20 split shape/expression rows, 18 x 32
#ifndef K2B_TC_EPI_WARPS
#define K2B_TC_EPI_WARPS 16
#endif
constexpr int kTcEpiWarps = K2B_TC_EPI_WARPS;          // 8 | 16: (lane quarter) x (tile of the pair) x (frame part)
constexpr int kTcEpiParts = kTcEpiWarps / 8;           // warps sharing one accumulator quarter, each takes FR / parts frames
constexpr int kTcThreads = 32 * (kTcEpiWarps + 2);     // epilogue warps, then the TMA producer warp, then the MMA issuer
constexpr int kTcStages = 4;       // ring depth (K blocks of dirs, 16 KB each) of the plain blend
constexpr int kTcStagesFused = 6;
constexpr int kTcFusedJoints = 24; // the fused blend + skinning kernel is instantiated for the SMPL skeleton
constexpr int kTcAccStages = 4;    // accumulators in TMEM (two pairs)
constexpr int kTcTmemCols = 512;

__host__ __device__ constexpr int tc_kpad(int kdepth) { return (kdepth + kTcBK - 1) / kTcBK * kTcBK; }
__host__ __device__ constexpr int tc_b_bytes() { return kTcN * 128; }                    // one (column tile, K block)
__host__ __device__ constexpr int tc_f_bytes(int kpad, int fr) { return (kpad / kTcBK) * fr * 128; }
__host__ __device__ constexpr size_t tc_smem_bytes(int kpad, int fr, int stages, int fused_joints) {
  return (size_t)tc_f_bytes(kpad, fr) + (size_t)stages * (size_t)tc_b_bytes() + (size_t)fr * fused_joints * 48 + (fused_joints ? (size_t)fr * 16 : 0) + 1024;
}
// offset (in FP16 elements) of element (row, k < 64) inside one swizzle-128B K block
__host__ __device__ constexpr int tc_elem_off(int row, int k) {
  return (row * 128 + (((k / 8) ^ (row % 8)) * 16) + (k % 8) * 2) / 2;
}

#if defined(__CUDACC__)
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, FP32 accumulate
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset: unused for swizzled K-major (1)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;        // stride byte offset: 8 rows x 128 B, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                              // layout type 2 = SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): F16 x F16 -> F32 (A/B format 0), K-major A and B.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace tc

struct BlendParams {
  const float* feat;        // [frames_padded][npose] pose features (frames padded to 128, zero rows)
  const float* shape;       // [B][ns]
  const __half* b_tiles;    // [n_tiles (even)][kpad/64][tc_b_bytes/2] pre-tiled FP16 (dirs * dir_scale)^T blocks
  float inv_scale;          // 1 / dir_scale (a power of two)
  const float* v_template;  // [3V]
  const float4* skin;       // [frames_padded][nj][3] rows of the 3x4 skinning matrices (mesh_pose_kernel)
  const float* transl;      // [B][3] or null
  const int* ell_idx;       // [ell][V] joints of each vertex's non-zero skinning weights
  const float* ell_w;       // [ell][V]
  float* out;               // [B][3V] skinned vertices
  long num_frames;
  int npose, ns, kpad, nv, nj, ell, n_tiles;
  int debug;   // diagnostic builds only (-DK2B_DIAG, K2B_TC_DEBUG): 1 = skip output stores, 3 = skip the epilogue.
               // The shipped library compiles the switch out (K2B_TC_DBG is the constant 0).
};

// Column tiles carry VERTICES, not raw columns: tile row n = 32 q + l holds coordinate l % 3 of vertex
// tile * 40 + q * 10 + l / 3 (l < 30; rows 30, 31 of every quarter are zero padding), so the three
// coordinates of a vertex sit in neighbouring TMEM lanes of the same epilogue warp.
constexpr int kTcVertsPerQuarter = 10;
constexpr int kTcVertsPerTile = 4 * kTcVertsPerQuarter;
__host__ __device__ constexpr int tc_tile_vertex(int tile, int n) {
  return (n % 32) < 30 ? tile * kTcVertsPerTile + (n / 32) * kTcVertsPerQuarter + (n % 32) / 3 : -1;
}

// FR = frames per pass (MMA N), STAGES = dirs ring depth; NE > 0 (ELL width) with NJ joints = LBS skinning in the epilogue (the pass's
// skinning matrices are staged in shared memory; column tiles follow tc_tile_vertex) or plain v_posed output
// (column tile t = columns [128 t, 128 t + 128), skinned afterwards by skin_inplace_kernel)
template <int FR, int STAGES, int NE, int NJ>
__global__ void __launch_bounds__(kTcThreads, 1) blend_skin_tc_kernel(const __grid_constant__ BlendParams p) {
  constexpr bool FUSED = NE > 0;   // NE = skinning weights per vertex (ELL width, 1..4); 0 = unfused
  static_assert(FR % (16 * kTcEpiParts) == 0 && FR >= 16 && FR * kTcAccStages <= kTcTmemCols && STAGES <= 6, "tile shape");
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int kpad = p.kpad;
  constexpr int b_bytes = tc_b_bytes();
  const int f_bytes = tc_f_bytes(kpad, FR);
  __half* sF = reinterpret_cast<__half*>(tc_smem);                       // features of the pass
  unsigned char* sR = tc_smem + f_bytes;                                 // dirs ring
  float4* sA = reinterpret_cast<float4*>(tc_smem + f_bytes + (size_t)STAGES * b_bytes);   // FUSED: [FR][nj][3] rows
  const size_t a_bytes = FUSED ? (size_t)FR * NJ * 48 + (size_t)FR * 16 : 0;
  float4* sT = sA + FR * 3 * NJ;                                      // FUSED: [FR] translations
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + f_bytes + (size_t)STAGES * b_bytes + a_bytes);
  // bars: [0,6) ring_full; [6,12) ring_empty; [12,16) acc_full; [16,20) acc_empty; then the TMEM base word
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = tc::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(BAR(i), 1);          // ring_full: producer's expect_tx arrival
      tc::mbar_init(BAR(6 + i), 1);      // ring_empty: tcgen05.commit
    }
    for (int i = 0; i < kTcAccStages; ++i) {
      tc::mbar_init(BAR(12 + i), 1);     // acc_full: tcgen05.commit
      tc::mbar_init(BAR(16 + i), 128 * kTcEpiParts);   // acc_empty: the epilogue warps that drained it
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_word), kTcTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_word;

  const long num_passes = (p.num_frames + FR - 1) / FR;
  const uint32_t idesc = tc::make_idesc(kTcN, FR);       // M = columns tile, N = frames
  const int kblocks = kpad / kTcBK;
  const int npairs = p.n_tiles / 2;                         // n_tiles is padded to an even count
  uint32_t ph_rfull = 0, ph_rempty = 0, ph_afull = 0, ph_aempty = 0;   // one parity bit per stage
  long pair_seq = 0;   // running pair counter across passes (accumulator ring position)
  long blk_seq = 0;    // running K-block counter across passes (dirs ring position)

  for (long ps = blockIdx.x; ps < num_passes; ps += gridDim.x) {
    const long f0 = ps * FR;
    // every MMA of the previous pass has retired (the epilogue saw its last acc_full) before the
    // features are overwritten
    __syncthreads();
    // ---- N operand: features of 128 frames -> shared memory (canonical K-major image) ----------
    for (int i = tid; i < FR * kpad; i += kTcThreads) {
      const int r = i / kpad, k = i - r * kpad;
      const long f = f0 + r;
      __half x = __float2half_rn(0.f);
      if (k < p.npose) {
        x = __float2half_rn(p.feat[f * p.npose + k]);       // feat rows are padded to a multiple of 128
      } else if (k < p.npose + 3 * p.ns && f < p.num_frames) {
        const int part = (k - p.npose) / p.ns, s = (k - p.npose) - part * p.ns;
        const float b = p.shape[f * p.ns + s];
        const __half hi = __float2half_rn(b);
        x = part == 1 ? __float2half_rn(b - __half2float(hi)) : hi;   // [hi | lo | hi]
      }
      sF[(k >> 6) * (FR * 64) + tc_elem_off(r, k & 63)] = x;
    }
    if constexpr (FUSED) {     // this pass's skinning matrices (rows past the last frame are never read)
      // The frame's translation is folded into the last column of every joint's matrix: the skinning weights of a
      // vertex sum to one, so sum_k w_k (A_k.w + t) = sum_k w_k A_k.w + t -- one shared-memory read less per output.
      const float4* src = p.skin + f0 * (3L * NJ);
      for (int i = tid; i < FR * 3 * NJ; i += kTcThreads) {
        float4 r = src[i];
        const int fr = i / (3 * NJ), c = i % 3;
        const long f = f0 + fr;
        if (p.transl && f < p.num_frames) r.w += p.transl[f * 3 + c];
        sA[i] = r;
      }
    }
    tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();

    if (warp == kTcEpiWarps) {
      // ---- TMA producer: blocks in the order the MMA warp consumes them: (2p,kb), (2p+1,kb) -------
      if (lane == 0) {
        long seq = blk_seq;
        for (int pr = 0; pr < npairs; ++pr)
          for (int kb = 0; kb < kblocks; ++kb)
            for (int h = 0; h < 2; ++h, ++seq) {
              const int s = (int)(seq % STAGES);
              if (seq >= STAGES) {
                tc::mbar_wait(BAR(6 + s), (ph_rempty >> s) & 1u);
                ph_rempty ^= 1u << s;
              }
              const size_t blk = (size_t)(2 * pr + h) * kblocks + kb;
              tc::mbar_expect_tx(BAR(s), (uint32_t)b_bytes);
              tc::bulk_g2s(tc::smem_u32(sR + (size_t)s * b_bytes), p.b_tiles + blk * (b_bytes / 2), (uint32_t)b_bytes,
                           BAR(s));
            }
      }
    } else if (warp == kTcEpiWarps + 1) {
      // ---- MMA issuer: one thread; two interleaved accumulation chains per pair -------------------
      if (lane == 0) {
        const uint32_t f_addr = tc::smem_u32(sF);
        long seq = blk_seq;
        for (int pr = 0; pr < npairs; ++pr) {
          const long pseq = pair_seq + pr;
          const int a0 = (int)(pseq & 1) * 2;
          if (pseq >= 2) {
            tc::mbar_wait(BAR(16 + a0), (ph_aempty >> a0) & 1u);
            ph_aempty ^= 1u << a0;
            tc::mbar_wait(BAR(16 + a0 + 1), (ph_aempty >> (a0 + 1)) & 1u);
            ph_aempty ^= 1u << (a0 + 1);
          }
          const uint32_t d0 = tmem_base + (uint32_t)(a0 * FR), d1 = d0 + (uint32_t)FR;
          for (int kb = 0; kb < kblocks; ++kb, seq += 2) {
            const int s0 = (int)(seq % STAGES), s1 = (int)((seq + 1) % STAGES);
            tc::mbar_wait(BAR(s0), (ph_rfull >> s0) & 1u);
            ph_rfull ^= 1u << s0;
            tc::mbar_wait(BAR(s1), (ph_rfull >> s1) & 1u);
            ph_rfull ^= 1u << s1;
            tc::tc_fence_after();
            const uint32_t r0 = tc::smem_u32(sR + (size_t)s0 * b_bytes), r1 = tc::smem_u32(sR + (size_t)s1 * b_bytes);
#pragma unroll
            for (int j = 0; j < kTcBK / 16; ++j) {
              const uint64_t fdesc = tc::make_desc(f_addr + (uint32_t)(kb * (FR * 128) + j * 32));
              const uint32_t acc = (kb | j) ? 1u : 0u;
              tc::mma_f16(d0, tc::make_desc(r0 + j * 32), fdesc, idesc, acc);
              tc::mma_f16(d1, tc::make_desc(r1 + j * 32), fdesc, idesc, acc);
            }
            tc::mma_commit(BAR(6 + s0));    // ring stages free once these MMAs retire
            tc::mma_commit(BAR(6 + s1));
          }
          tc::mma_commit(BAR(12 + a0));     // both accumulators of the pair ready
          tc::mma_commit(BAR(12 + a0 + 1));
        }
      }
    } else {
      if constexpr (FUSED) {
        // ---- epilogue warps 0-7: lane quarter q = warp % 4, tile h = warp / 4 of every pair -----------
        // Each lane owns ONE output coordinate c3 of one vertex: it fetches the vertex's blended position
        // from its two neighbour lanes, builds row c3 of the vertex's skinning matrix
        // sum_k w_k A[frame][joint_k] and writes verts[frame][vertex][c3] -- v_posed never leaves the SM.
        // All per-frame strides are compile-time (NJ) or running pointers, so a frame costs ~25 instructions.
        const int q = warp & 3, h = (warp >> 2) & 1, part = warp >> 3;
        constexpr int CH = FR / 16 / kTcEpiParts;      // 16-frame chunks per warp
        const int c3 = lane % 3, lbase = lane - c3;
        // the two other coordinates of this lane's vertex come from lanes lbase + (c3 + 1) % 3 and lbase + (c3 + 2) % 3
        // (two shuffles; the matrix row is rotated to match instead of fetching x, y, z with three)
        const int s1 = lbase + (c3 + 1) % 3 < 32 ? lbase + (c3 + 1) % 3 : 31;
        const int s2 = lbase + (c3 + 2) % 3 < 32 ? lbase + (c3 + 2) % 3 : 31;
        const long ncols = 3L * p.nv;
        for (int pr = 0; pr < npairs; ++pr) {
          const long pseq = pair_seq + pr;
          const int a = (int)(pseq & 1) * 2 + h;
          const int v = tc_tile_vertex(2 * pr + h, q * 32 + lane);
          const bool v_ok = v >= 0 && v < p.nv;
          const int col = v_ok ? 3 * v + c3 : 0;
          const float tv = v_ok ? __ldg(p.v_template + col) : 0.f;   // issued before the wait
          const float4* Ak[NE];
          float wk[NE];
#pragma unroll
          for (int k = 0; k < NE; ++k) {
            Ak[k] = sA + 3 * (v_ok ? __ldg(p.ell_idx + (long)k * p.nv + v) : 0) + c3;
            wk[k] = v_ok ? __ldg(p.ell_w + (long)k * p.nv + v) : 0.f;
          }
          tc::mbar_wait(BAR(12 + a), (ph_afull >> a) & 1u);
          ph_afull ^= 1u << a;
          tc::tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * FR);
          const bool st_ok = v_ok && K2B_TC_DBG(p) != 1;
          if (K2B_TC_DBG(p) == 3) {                                      // timing experiment: no TMEM reads / stores
            tc::tc_fence_before();
            tc::mbar_arrive(BAR(16 + a));
            continue;
          }
          float* o = p.out + (f0 + part * (CH * 16)) * ncols + col;
          const long left = p.num_frames - f0;
          int nvalid = (st_ok ? (left < FR ? (int)left : FR) : 0) - part * (CH * 16);   // frames this lane stores
#pragma unroll 1
          for (int ch = part * CH; ch < (part + 1) * CH; ++ch) {
            uint32_t acc[16];
            tc::tmem_ld16(taddr + (uint32_t)(ch * 16), acc);
            tc::tmem_ld_wait();
            if (ch == (part + 1) * CH - 1) {
              tc::tc_fence_before();
              tc::mbar_arrive(BAR(16 + a));                        // accumulator may be overwritten
            }
            constexpr int FS = 3 * NJ;                             // float4 rows per frame
#pragma unroll
            for (int i = 0; i < 16; ++i) {                         // fully unrolled, branch-free: 16 frames overlap
              const float pc = fmaf(__uint_as_float(acc[i]), p.inv_scale, tv);
              const float p1 = __shfl_sync(0xffffffffu, pc, s1), p2 = __shfl_sync(0xffffffffu, pc, s2);
              float4 T = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int k = 0; k < NE; ++k) {
                const float4 r = Ak[k][(ch * 16 + i) * FS];
                T.x = fmaf(wk[k], r.x, T.x); T.y = fmaf(wk[k], r.y, T.y);
                T.z = fmaf(wk[k], r.z, T.z); T.w = fmaf(wk[k], r.w, T.w);
              }
              // row entries for (own, next, next-next) coordinate: c3 = 0 -> (x, y, z), 1 -> (y, z, x), 2 -> (z, x, y)
              const float t0 = c3 == 0 ? T.x : (c3 == 1 ? T.y : T.z);
              const float t1 = c3 == 0 ? T.y : (c3 == 1 ? T.z : T.x);
              const float t2 = c3 == 0 ? T.z : (c3 == 1 ? T.x : T.y);
              const float ov = fmaf(t0, pc, fmaf(t1, p1, fmaf(t2, p2, T.w)));
              if (i < nvalid) o[0] = ov;
              o += ncols;
            }
            nvalid -= 16;
          }
        }
      } else {
        // ---- epilogue warps 0-7: lane quarter q = warp % 4, tile h = warp / 4 of every pair -----------
        const int q = warp & 3, h = (warp >> 2) & 1, part = warp >> 3;
        constexpr int CH = FR / 16 / kTcEpiParts;
        for (int pr = 0; pr < npairs; ++pr) {
          const long pseq = pair_seq + pr;
          const int a = (int)(pseq & 1) * 2 + h;
          const int c = (2 * pr + h) * kTcN + q * 32 + lane;       // this lane's output column
          const bool col_ok = c < 3 * p.nv;
          const float tv = col_ok ? __ldg(p.v_template + c) : 0.f; // issued before the wait
          tc::mbar_wait(BAR(12 + a), (ph_afull >> a) & 1u);
          ph_afull ^= 1u << a;
          tc::tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * FR);
          const long ncols = 3L * p.nv;
          float* o = p.out + f0 * ncols + c;
          if (K2B_TC_DBG(p) == 3) {                                      // timing experiment: no TMEM reads / stores
            tc::tc_fence_before();
            tc::mbar_arrive(BAR(16 + a));
            continue;
          }
  #pragma unroll 1
          for (int ch = part * CH; ch < (part + 1) * CH; ++ch) {
            uint32_t v[16];
            tc::tmem_ld16(taddr + (uint32_t)(ch * 16), v);
            tc::tmem_ld_wait();
            if (ch == (part + 1) * CH - 1) {
              tc::tc_fence_before();
              tc::mbar_arrive(BAR(16 + a));                        // accumulator may be overwritten
            }
            if (col_ok && K2B_TC_DBG(p) != 1) {
  #pragma unroll
              for (int i = 0; i < 16; ++i) {
                const long fr = ch * 16 + i;
                if (f0 + fr < p.num_frames) o[fr * ncols] = fmaf(__uint_as_float(v[i]), p.inv_scale, tv);
              }
            }
          }
        }
      }
    }
    pair_seq += npairs;
    blk_seq += (long)p.n_tiles * kblocks;
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, kTcTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused blend + skinning with lane = VERTEX (round 2).  A column tile is 128 vertices; it has THREE accumulators
// D_c[vertex lane][frame], c = x, y, z (three MMA chains over the same frame operand, 3 x FR TMEM columns), so a lane
// reads its vertex's three coordinates straight from tensor memory: no shuffles, every lane useful (the
// lane = coordinate layout above uses 30 of 32 and needs two shuffles per output), and the three rows of a joint's
// skinning matrix are loaded once per vertex instead of once per coordinate lane: 2 NE x 3 LDS.128 per 96 outputs
// = 0.25 shared-memory wavefronts per output instead of 0.43 -- the shared-memory pipe is what bounds the kernel
// (profiles/r02_blend_skin_ncu_metrics.txt: 81 % of peak).  Price: a lane stores x, y, z of its vertex (12 bytes), so a
// store instruction writes 4 bytes every 12 (three instructions fill the 384-byte run).  Two tiles are in flight (two
// accumulator stages of 3 x FR columns); dirs blocks stream through the same TMA ring, three per K block.
// b_tiles layout: [tile][coordinate][K block][128 rows x 128 B swizzled image], row n = vertex 128 tile + n.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTcVtVerts = 128;      // vertices per tile of the lane = vertex kernel
constexpr int kTcVtAccStages = 2;

template <int FR, int STAGES, int NE, int NJ>
__global__ void __launch_bounds__(kTcThreads, 1) blend_skin_vt_kernel(const __grid_constant__ BlendParams p) {
  static_assert(FR % (16 * kTcEpiParts) == 0 && 3 * FR * kTcVtAccStages <= kTcTmemCols && STAGES <= 6 && STAGES % 3 == 0,
                "tile shape");
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int kpad = p.kpad;
  constexpr int b_bytes = tc_b_bytes();
  const int f_bytes = tc_f_bytes(kpad, FR);
  __half* sF = reinterpret_cast<__half*>(tc_smem);
  unsigned char* sR = tc_smem + f_bytes;
  float4* sA = reinterpret_cast<float4*>(tc_smem + f_bytes + (size_t)STAGES * b_bytes);   // [FR][NJ][3] rows, translation folded in
  const size_t a_bytes = (size_t)FR * NJ * 48 + (size_t)FR * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + f_bytes + (size_t)STAGES * b_bytes + a_bytes);
  // bars: [0,6) ring_full; [6,12) ring_empty; [12,14) acc_full; [16,18) acc_empty; then the TMEM base word
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = tc::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(BAR(i), 1);
      tc::mbar_init(BAR(6 + i), 1);
    }
    for (int i = 0; i < kTcVtAccStages; ++i) {
      tc::mbar_init(BAR(12 + i), 1);                       // acc_full: tcgen05.commit
      tc::mbar_init(BAR(16 + i), 128 * kTcEpiParts);       // acc_empty: the epilogue warps that drained the stage
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_word), kTcTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_word;

  const long num_passes = (p.num_frames + FR - 1) / FR;
  const uint32_t idesc = tc::make_idesc(kTcN, FR);
  const int kblocks = kpad / kTcBK;
  const int ntiles = p.n_tiles;
  uint32_t ph_rfull = 0, ph_rempty = 0, ph_afull = 0, ph_aempty = 0;
  long tile_seq = 0;   // running tile counter across passes (accumulator stage = tile_seq & 1)
  long blk_seq = 0;    // running dirs-block counter across passes (ring position)

  for (long ps = blockIdx.x; ps < num_passes; ps += gridDim.x) {
    const long f0 = ps * FR;
    __syncthreads();      // every MMA of the previous pass has retired before the features are overwritten
    for (int i = tid; i < FR * kpad; i += kTcThreads) {
      const int r = i / kpad, k = i - r * kpad;
      const long f = f0 + r;
      __half x = __float2half_rn(0.f);
      if (k < p.npose) {
        x = __float2half_rn(p.feat[f * p.npose + k]);
      } else if (k < p.npose + 3 * p.ns && f < p.num_frames) {
        const int part = (k - p.npose) / p.ns, s = (k - p.npose) - part * p.ns;
        const float b = p.shape[f * p.ns + s];
        const __half hi = __float2half_rn(b);
        x = part == 1 ? __float2half_rn(b - __half2float(hi)) : hi;   // [hi | lo | hi]
      }
      sF[(k >> 6) * (FR * 64) + tc_elem_off(r, k & 63)] = x;
    }
    {
      const float4* src = p.skin + f0 * (3L * NJ);
      for (int i = tid; i < FR * 3 * NJ; i += kTcThreads) {
        float4 r = src[i];
        const int fr = i / (3 * NJ), c = i % 3;
        const long f = f0 + fr;
        if (p.transl && f < p.num_frames) r.w += p.transl[f * 3 + c];    // sum_k w_k = 1: the translation rides along
        sA[i] = r;
      }
    }
    tc::fence_proxy_async();
    __syncthreads();

    if (warp == kTcEpiWarps) {
      // ---- TMA producer: blocks in the order the MMA warp consumes them: (tile, kb, coordinate) -----------------
      if (lane == 0) {
        long seq = blk_seq;
        for (int t = 0; t < ntiles; ++t)
          for (int kb = 0; kb < kblocks; ++kb)
            for (int c = 0; c < 3; ++c, ++seq) {
              const int s = (int)(seq % STAGES);
              if (seq >= STAGES) {
                tc::mbar_wait(BAR(6 + s), (ph_rempty >> s) & 1u);
                ph_rempty ^= 1u << s;
              }
              const size_t blk = ((size_t)t * 3 + c) * kblocks + kb;
              tc::mbar_expect_tx(BAR(s), (uint32_t)b_bytes);
              tc::bulk_g2s(tc::smem_u32(sR + (size_t)s * b_bytes), p.b_tiles + blk * (b_bytes / 2), (uint32_t)b_bytes, BAR(s));
            }
      }
    } else if (warp == kTcEpiWarps + 1) {
      // ---- MMA issuer: three interleaved accumulation chains (x, y, z) per tile ------------------------------------
      if (lane == 0) {
        const uint32_t f_addr = tc::smem_u32(sF);
        long seq = blk_seq;
        for (int t = 0; t < ntiles; ++t) {
          const long ts = tile_seq + t;
          const int a = (int)(ts & 1);
          if (ts >= kTcVtAccStages) {
            tc::mbar_wait(BAR(16 + a), (ph_aempty >> a) & 1u);
            ph_aempty ^= 1u << a;
          }
          const uint32_t d0 = tmem_base + (uint32_t)(a * 3 * FR);
          for (int kb = 0; kb < kblocks; ++kb, seq += 3) {
            uint32_t rr[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int s = (int)((seq + c) % STAGES);
              tc::mbar_wait(BAR(s), (ph_rfull >> s) & 1u);
              ph_rfull ^= 1u << s;
              rr[c] = tc::smem_u32(sR + (size_t)s * b_bytes);
            }
            tc::tc_fence_after();
#pragma unroll
            for (int j = 0; j < kTcBK / 16; ++j) {
              const uint64_t fdesc = tc::make_desc(f_addr + (uint32_t)(kb * (FR * 128) + j * 32));
              const uint32_t acc = (kb | j) ? 1u : 0u;
#pragma unroll
              for (int c = 0; c < 3; ++c) tc::mma_f16(d0 + (uint32_t)(c * FR), tc::make_desc(rr[c] + j * 32), fdesc, idesc, acc);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) tc::mma_commit(BAR(6 + (int)((seq + c) % STAGES)));
          }
          tc::mma_commit(BAR(12 + a));
        }
      }
    } else {
      // ---- epilogue warps: lane quarter q = warp % 4, accumulator stage h = tile parity, frame part ----------------
      const int q = warp & 3, h = (warp >> 2) & 1, part = warp >> 3;
      constexpr int CH = FR / 16 / kTcEpiParts;
      constexpr int FS = 3 * NJ;                              // float4 rows per frame
      const long ncols = 3L * p.nv;
      for (int t = 0; t < ntiles; ++t) {
        const long ts = tile_seq + t;
        if ((int)(ts & 1) != h) continue;
        const int v = t * kTcVtVerts + q * 32 + lane;
        const bool v_ok = v < p.nv;
        float tv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) tv[c] = v_ok ? __ldg(p.v_template + 3 * v + c) : 0.f;
        const float4* Ak[NE];
        float wk[NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) {
          Ak[k] = sA + 3 * (v_ok ? __ldg(p.ell_idx + (long)k * p.nv + v) : 0);
          wk[k] = v_ok ? __ldg(p.ell_w + (long)k * p.nv + v) : 0.f;
        }
        tc::mbar_wait(BAR(12 + h), (ph_afull >> h) & 1u);
        ph_afull ^= 1u << h;
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 3 * FR);
        float* o = p.out + (f0 + part * (CH * 16)) * ncols + 3L * v;
        const long left = p.num_frames - f0;
        int nvalid = (v_ok ? (left < FR ? (int)left : FR) : 0) - part * (CH * 16);
#pragma unroll 1
        for (int ch = part * CH; ch < (part + 1) * CH; ++ch) {
          uint32_t ax[16], ay[16], az[16];
          tc::tmem_ld16(taddr + (uint32_t)(ch * 16), ax);
          tc::tmem_ld16(taddr + (uint32_t)(FR + ch * 16), ay);
          tc::tmem_ld16(taddr + (uint32_t)(2 * FR + ch * 16), az);
          tc::tmem_ld_wait();
          if (ch == (part + 1) * CH - 1) {
            tc::tc_fence_before();
            tc::mbar_arrive(BAR(16 + h));                      // the stage may be overwritten
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float px = fmaf(__uint_as_float(ax[i]), p.inv_scale, tv[0]);
            const float py = fmaf(__uint_as_float(ay[i]), p.inv_scale, tv[1]);
            const float pz = fmaf(__uint_as_float(az[i]), p.inv_scale, tv[2]);
            float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
#pragma unroll
            for (int k = 0; k < NE; ++k) {
              const float4* r = Ak[k] + (ch * 16 + i) * FS;
              const float4 r0 = r[0], r1 = r[1], r2 = r[2];
              T0.x = fmaf(wk[k], r0.x, T0.x); T0.y = fmaf(wk[k], r0.y, T0.y); T0.z = fmaf(wk[k], r0.z, T0.z); T0.w = fmaf(wk[k], r0.w, T0.w);
              T1.x = fmaf(wk[k], r1.x, T1.x); T1.y = fmaf(wk[k], r1.y, T1.y); T1.z = fmaf(wk[k], r1.z, T1.z); T1.w = fmaf(wk[k], r1.w, T1.w);
              T2.x = fmaf(wk[k], r2.x, T2.x); T2.y = fmaf(wk[k], r2.y, T2.y); T2.z = fmaf(wk[k], r2.z, T2.z); T2.w = fmaf(wk[k], r2.w, T2.w);
            }
            if (i < nvalid) {
              // streaming stores (evict-first): 83 KB per frame that nothing on the GPU reads again must not push the fit
              // kernel's L2-resident L-BFGS history (it runs beside this kernel) out of the cache
              __stcs(o + 0, fmaf(T0.x, px, fmaf(T0.y, py, fmaf(T0.z, pz, T0.w))));
              __stcs(o + 1, fmaf(T1.x, px, fmaf(T1.y, py, fmaf(T1.z, pz, T1.w))));
              __stcs(o + 2, fmaf(T2.x, px, fmaf(T2.y, py, fmaf(T2.z, pz, T2.w))));
            }
            o += ncols;
          }
          nvalid -= 16;
        }
      }
    }
    tile_seq += ntiles;
    blk_seq += (long)ntiles * 3 * kblocks;
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, kTcTmemCols);
}
#endif  // __CUDACC__

}  // namespace k2b
