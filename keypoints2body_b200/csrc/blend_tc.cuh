// blend_tc.cuh -- K2: pose / shape blend as a TF32 GEMM on the 5th-gen tensor cores (tcgen05).
//
//   v_posed[f][c] = v_template[c] + sum_k feat[f][k] * dirs[k][c],   c in [0, 3V)
//   feat[f] = [ (R_1..R_{nj-1} - I).flatten() (9(nj-1)) | shape_hi | shape_lo | shape_hi | 0-pad ]
//   dirs    = [ posedirs ; S_hi ; S_hi ; S_lo ]
// The shape blend carries decimetre-scale offsets, so its 10 rows are split into TF32 hi/lo parts on
// both sides (x = x_hi + x_lo, x_hi*S_hi + x_lo*S_hi + x_hi*S_lo; only the lo*lo term is dropped):
// that keeps the blend within ~1e-6 m of FP32 instead of ~7e-5 m for plain TF32.
//
// smplx does this as `pose_feature @ posedirs` + `blend_shapes` [smplx-from-memory]; the reference
// reaches it through the final forward (/root/reference/keypoints2body/core/fitters/world_space.py:258-278).
//
// Tiling: M = 128 frames per CTA pass.  The A operand (the frames' features, 128 x Kpad TF32)
// lives in TENSOR MEMORY (Kpad <= 240 columns, written once per pass with tcgen05.st), so shared
// memory is free for a 6-deep ring of B blocks.  An output tile is N = 128 columns (wide MMAs: a
// chain of MMAs into one accumulator is latency-bound, ~130 cycles each, so N must be large
// enough that the math time, N/2 cycles, is comparable); K is walked in blocks of 48 (6 MMAs of
// K = 8).  The B operand (dirs) is pre-tiled on the host into the exact shared-memory image of each
// (N tile, K block), so one elected thread streams a block with ONE 1-D TMA bulk copy
// (cp.async.bulk -> mbarrier complete_tx).  One thread issues tcgen05.mma (A from TMEM, B from
// smem descriptors, accumulators in TMEM: 2 stages x 128 columns); eight epilogue warps (row
// quarter x column half) drain TMEM with tcgen05.ld, transpose 32x32 blocks through a padded
// staging tile and write fp32 rows as full 128-byte segments (+ v_template).
//
// B block layout (no swizzle, K-major "interleaved" canonical UMMA layout, cute mma_traits_sm100):
//   byte address(row r, k) = (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4
// i.e. 8-row x 16-byte core matrices, contiguous along K (LBO = 128 B), SBO = (48 / 4) * 128 B.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace k2b {

constexpr int kTcM = 128;          // frames per pass
constexpr int kTcN = 128;          // output columns per tile
constexpr int kTcBK = 48;          // K block streamed per TMA copy (6 MMAs of K = 8)
constexpr int kTcKpadMax = 240;    // SMPL: 207 pose features + 3 x 10 split shape rows, padded to 5 x 48
constexpr int kTcThreads = 320;    // warps 0-7 epilogue, warp 8 TMA producer, warp 9 MMA issuer
constexpr int kTcStages = 6;       // B ring depth (K blocks)
constexpr int kTcAccStages = 2;    // accumulator stages in TMEM
constexpr int kTcTmemCols = 512;   // A: columns [0,240); accumulators at columns 256 and 384
constexpr int kTcStageStride = 33; // padded row stride (floats) of the per-warp epilogue staging tile

__host__ __device__ constexpr int tc_kpad(int kdepth) { return (kdepth + kTcBK - 1) / kTcBK * kTcBK; }
__host__ __device__ constexpr int tc_sbo_bytes() { return (kTcBK / 4) * 128; }
__host__ __device__ constexpr int tc_b_bytes() { return (kTcN / 8) * tc_sbo_bytes(); }   // one (N tile, K block)
__host__ __device__ constexpr int tc_stage_bytes() { return 8 * 32 * kTcStageStride * 4; }
__host__ __device__ constexpr size_t tc_smem_bytes() {
  return (size_t)kTcStages * (size_t)tc_b_bytes() + tc_stage_bytes() + 1024;
}
// offset (in floats) of element (row, k) inside a B block image
__host__ __device__ constexpr int tc_elem_off(int row, int k) {
  return ((row / 8) * tc_sbo_bytes() + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4) / 4;
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
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;         // leading byte offset (K direction), bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride byte offset (8-row groups), bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
  return d;                                            // base offset 0, layout type 0 = SWIZZLE_NONE
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): TF32 x TF32 -> F32, K-major A and B.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace tc

struct BlendParams {
  const float* feat;        // [frames_padded][npose] pose features (frames padded to kTcM, zero rows)
  const float* shape;       // [B][ns]
  const float* b_tiles;     // [n_tiles][kpad/48][tc_b_bytes/4] pre-tiled TF32 dirs
  const float* v_template;  // [ncols]
  float* out;               // [B][ncols]  (ncols = 3V)
  long num_frames;
  int npose, ns, kpad, ncols, n_tiles;
  int debug;   // K2B_TC_DEBUG: 1 = skip output stores, 2 = skip B copies (timing experiments only)
};

__global__ void __launch_bounds__(kTcThreads, 1) blend_tc_kernel(const __grid_constant__ BlendParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int kpad = p.kpad;
  constexpr int b_bytes = tc_b_bytes(), sbo = tc_sbo_bytes();
  unsigned char* sB = tc_smem;
  float* sStage = reinterpret_cast<float*>(tc_smem + (size_t)kTcStages * b_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + (size_t)kTcStages * b_bytes + tc_stage_bytes());
  // bars: [0,6) b_full; [6,12) b_empty; [12,14) acc_full; [14,16) acc_empty; then the TMEM base word
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = tc::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (tid == 0) {
    for (int i = 0; i < kTcStages; ++i) {
      tc::mbar_init(BAR(i), 1);          // b_full: producer's expect_tx arrival
      tc::mbar_init(BAR(6 + i), 1);      // b_empty: tcgen05.commit
    }
    for (int i = 0; i < kTcAccStages; ++i) {
      tc::mbar_init(BAR(12 + i), 1);     // acc_full: tcgen05.commit
      tc::mbar_init(BAR(14 + i), 256);   // acc_empty: all eight epilogue warps
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_word), kTcTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_word;
  const uint32_t tmem_acc = tmem_base + 256u;                 // accumulators: columns 256.. and 384..

  const long num_mtiles = (p.num_frames + kTcM - 1) / kTcM;
  const uint32_t idesc = tc::make_idesc(kTcM, kTcN);
  const int kblocks = kpad / kTcBK;
  uint32_t ph_bfull = 0, ph_bempty = 0, ph_afull = 0, ph_aempty = 0;   // one parity bit per stage
  long tile_seq = 0;   // running N-tile counter across passes (accumulator ring position)
  long blk_seq = 0;    // running K-block counter across passes (B ring position)

  for (long mt = blockIdx.x; mt < num_mtiles; mt += gridDim.x) {
    const long f0 = mt * kTcM;
    // every MMA of the previous pass has retired (the epilogue saw its last acc_full) before A is rewritten
    __syncthreads();
    // ---- A operand -> tensor memory: warps 0-3, thread = frame row, Kpad columns ------------------
    if (warp < 4) {
      const long f = f0 + warp * 32 + lane;
      const bool live = f < p.num_frames;
      const float* fr = p.feat + f * p.npose;               // feat rows are padded to a multiple of kTcM
      const float* sh = p.shape + (live ? f : 0) * p.ns;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int k0 = 0; k0 < kpad; k0 += 8) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = k0 + i;
          float x = 0.f;
          if (k < p.npose) {
            x = tc::to_tf32(fr[k]);
          } else if (k < p.npose + 3 * p.ns && live) {
            const int part = (k - p.npose) / p.ns, s = (k - p.npose) - part * p.ns;
            const float b = sh[s];
            const float hi = tc::to_tf32(b);
            x = part == 1 ? tc::to_tf32(b - hi) : hi;       // [hi | lo | hi]
          }
          v[i] = __float_as_uint(x);
        }
        tc::tmem_st8(trow + (uint32_t)k0, v);
      }
      tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    if (warp == 8) {
      // ---- TMA producer: one bulk copy per (N tile, K block) into a 6-deep ring --------------------
      if (lane == 0) {
        const int nblk = p.n_tiles * kblocks;
        for (int i = 0; i < nblk; ++i) {
          const long seq = blk_seq + i;
          const int s = (int)(seq % kTcStages);
          if (seq >= kTcStages) {
            tc::mbar_wait(BAR(6 + s), (ph_bempty >> s) & 1u);
            ph_bempty ^= 1u << s;
          }
          if (p.debug == 2 && seq >= kTcStages) {
            tc::mbar_arrive(BAR(s));
            continue;
          }
          tc::mbar_expect_tx(BAR(s), (uint32_t)b_bytes);
          tc::bulk_g2s(tc::smem_u32(sB + (size_t)s * b_bytes), p.b_tiles + (size_t)i * (b_bytes / 4), (uint32_t)b_bytes,
                       BAR(s));
        }
      }
    } else if (warp == 9) {
      // ---- MMA issuer: one thread; per tile Kpad/8 x (M128 N128 K8), A from TMEM --------------------
      if (lane == 0) {
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          const long tseq = tile_seq + nt;
          const int sa = (int)(tseq % kTcAccStages);
          if (tseq >= kTcAccStages) {
            tc::mbar_wait(BAR(14 + sa), (ph_aempty >> sa) & 1u);   // epilogue drained this accumulator
            ph_aempty ^= 1u << sa;
          }
          const uint32_t d_tmem = tmem_acc + (uint32_t)(sa * kTcN);
          for (int kb = 0; kb < kblocks; ++kb) {
            const long seq = blk_seq + (long)nt * kblocks + kb;
            const int s = (int)(seq % kTcStages);
            tc::mbar_wait(BAR(s), (ph_bfull >> s) & 1u);            // B block landed
            ph_bfull ^= 1u << s;
            tc::tc_fence_after();
            const uint32_t b_addr = tc::smem_u32(sB + (size_t)s * b_bytes);
#pragma unroll
            for (int j = 0; j < kTcBK / 8; ++j)
              tc::mma_tf32_ts(d_tmem, tmem_base + (uint32_t)(kb * kTcBK + j * 8), tc::make_desc(b_addr + j * 256, sbo),
                              idesc, (kb | j) ? 1u : 0u);
            tc::mma_commit(BAR(6 + s));     // B stage free once these MMAs retire
          }
          tc::mma_commit(BAR(12 + sa));     // accumulator ready
        }
      }
    } else {
      // ---- epilogue warps 0-7: row quarter q = warp % 4, column half h = warp / 4 ------------------
      const int q = warp & 3, hcol = warp >> 2;
      float* stg = sStage + warp * 32 * kTcStageStride;      // this warp's 32 x 32 block (stride 33)
      const int half = lane >> 4, cpair = (lane & 15) * 2;   // lane -> (row parity, column pair)
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        const long tseq = tile_seq + nt;
        const int sa = (int)(tseq % kTcAccStages);
        tc::mbar_wait(BAR(12 + sa), (ph_afull >> sa) & 1u);
        ph_afull ^= 1u << sa;
        tc::tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {                     // two 32-column chunks of this warp's half
          const int cc = hcol * 64 + ch * 32;                // column offset inside the tile
          const int c = nt * kTcN + cc + cpair;
          const float t0 = c < p.ncols ? __ldg(p.v_template + c) : 0.f;
          const float t1 = c + 1 < p.ncols ? __ldg(p.v_template + c + 1) : 0.f;
          const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(sa * kTcN + cc);
          uint32_t v[2][16];
          tc::tmem_ld16(taddr, v[0]);
          tc::tmem_ld16(taddr + 16, v[1]);
          tc::tmem_ld_wait();
          if (ch == 1) {
            tc::tc_fence_before();
            tc::mbar_arrive(BAR(14 + sa));                   // accumulator stage may be overwritten
          }
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 16; ++i) stg[lane * kTcStageStride + h * 16 + i] = __uint_as_float(v[h][i]);
          __syncwarp();
#pragma unroll 4
          for (int r2 = 0; r2 < 32; r2 += 2) {
            const int r = r2 + half;
            const long f = f0 + q * 32 + r;
            if (f < p.num_frames && p.debug != 1) {
              float2 w;
              w.x = stg[r * kTcStageStride + cpair] + t0;
              w.y = stg[r * kTcStageStride + cpair + 1] + t1;
              float* o = p.out + f * (long)p.ncols + c;
              if (c + 1 < p.ncols) *reinterpret_cast<float2*>(o) = w;
              else if (c < p.ncols) o[0] = w.x;
            }
          }
          __syncwarp();
        }
      }
    }
    tile_seq += p.n_tiles;
    blk_seq += (long)p.n_tiles * kblocks;
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, kTcTmemCols);
}
#endif  // __CUDACC__

}  // namespace k2b
