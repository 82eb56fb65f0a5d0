// Tensor-core recurrence kernels for large batches (B >> #SM, H = 64): K1c path_fwd_tc.
//
// At large batch the gate products of the recurrence are dense GEMMs: a CTA owns a tile of 128
// trajectories (the MMA M dimension, one TMEM lane per trajectory) and every step runs
//     D0[128,192]  = h0(t-1) . W_hh_l0^T                      (issued one step ahead)
//     D1[128,256]  = h1(t-1) . W_hh_l1^T  (+)=  h0(t) . W_ih_l1^T     (r, u share columns; n_i / n_h apart)
//     Dout[128,16] = h_top(t) . W_out^T
// as tcgen05.mma kind::f16 with FP32 accumulators in TMEM.  FP32-grade accuracy comes from a 3-pass
// split of both operands into fp16 hi + lo halves (hi*hi + hi*lo + lo*hi, 22 mantissa bits) after an
// exact power-of-two scaling that keeps both halves in fp16's normal range: weights by 2^a
// (max |w| 2^a in [2^13, 2^14)), hidden states (|h| < 1) by 2^14; the accumulator is scaled back by
// 2^-(a+14) in the epilogue.  fp16 rather than tf32 because the three recurrent matrices then fit in
// shared memory as resident hi/lo tiles (144 KB instead of 288 KB) and the MMA rate doubles.
//
// Warp roles (256 threads = 8 warps, 255 registers each): every warp is a gate-epilogue warp -- warp w reads TMEM
// lane quadrant w % 4 (32 trajectories) and hidden units [32 * (w / 4), +32): tcgen05.ld -> gates (raw MUFU) -> h(t)
// kept in registers for the next step's update, written as fp16 hi/lo into the K-major SWIZZLE_128B A tiles, and
// stashed for the backward.  Warp 0 also allocates TMEM and its lane 0 issues every MMA between its own epilogue
// phases (a dedicated ninth warp would cap the kernel at 168 registers per thread).
// mbarriers: d0 / d1 / out (tcgen05.commit: accumulators ready), a0 / a1 (256 arrivals: A tile written and
// accumulator drained), init (tile start).  The state-column and theta terms of layer 0 and the biases do not go
// through the tensor cores: theta and bias terms are folded into gi_ctx by K0 (per-trajectory row bias), the S state
// columns are S FMAs per gate in the epilogue.  gi_ctx and the stash use the row-fastest tiled layouts
// [tile][t][feature][128 rows]: a thread owns a trajectory row, so every warp access is one full 128-byte line.
#include "path_tc.cuh"

namespace visde {
namespace {

constexpr int kFwdThreads = 256;  // 8 warps = 255 registers per thread; warp 0 doubles as the MMA issuer

template <int NL, int S>
struct TcFwdSmem {
  static constexpr int NMAT = 2 * NL - 1;
  static constexpr int CS = (3 * S + 1 <= 4) ? 4 : (3 * S + 1 <= 8 ? 8 : 16);  // floats per unit in c0
  static constexpr int OFF_W = 0;                                  // [NMAT][hi, lo][192][128 B]
  static constexpr int OFF_WOUT = OFF_W + NMAT * 2 * kWTileBytes;  // [hi, lo][16][128 B]
  static constexpr int OFF_A = OFF_WOUT + 2 * kOutTileBytes;       // [NL][hi, lo][128][128 B]
  static constexpr int OFF_C0 = OFF_A + NL * 2 * kATileBytes;      // float [64][CS]: W_z rows, b_hn of layer 0
  static constexpr int OFF_C1 = OFF_C0 + 64 * CS * 4;              // float [64][4]: layer-1 bias terms
  static constexpr int OFF_OUTB = OFF_C1 + 64 * 4 * 4;             // float [16]
  static constexpr int OFF_BAR = OFF_OUTB + 64;
  struct Bars {
    uint64_t init, d0, a0, d1, a1, out;
    uint32_t tmem_base;
    uint32_t amax[2];
  };
  static constexpr size_t bytes = OFF_BAR + sizeof(Bars) + 1024;
};

// 12 MMAs: D[128, N] (+)= A[128, 64] . B[N, 64]^T with the 3-pass hi/lo split (small terms first)
__device__ __forceinline__ void issue_gemm(uint32_t dcol, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t dah = umma_desc(a_hi + j * 32, 16, 1024, 2), dal = umma_desc(a_lo + j * 32, 16, 1024, 2);
    const uint64_t dbh = umma_desc(b_hi + j * 32, 16, 1024, 2), dbl = umma_desc(b_lo + j * 32, 16, 1024, 2);
    umma_f16(dcol, dal, dbh, idesc, (accumulate || j > 0) ? 1u : 0u);
    umma_f16(dcol, dah, dbl, idesc, 1u);
    umma_f16(dcol, dah, dbh, idesc, 1u);
  }
}

template <int NL, int S>
__global__ void __launch_bounds__(kFwdThreads, 1) path_fwd_tc_kernel(PathParams p) {
  using L = TcFwdSmem<NL, S>;
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, CS = L::CS, NMAT = L::NMAT;
  static_assert(NOUT <= 16, "output projection tile holds 16 rows");
  constexpr uint32_t TMEM_COLS = NL == 2 ? 512 : 256;
  constexpr uint32_t D0_COL = 0, D1_COL = 192, DOUT_COL = NL == 2 ? 448 : 192;
  extern __shared__ __align__(1024) uint8_t smem_raw_tc[];
  // offset arithmetic on the __shared__ array (not a uintptr_t round trip) keeps the address space known: LDS / STS
  uint8_t* smem = smem_raw_tc + ((1024u - (smem_u32(smem_raw_tc) & 1023u)) & 1023u);
  typename L::Bars* bars = reinterpret_cast<typename L::Bars*>(smem + L::OFF_BAR);
  float* c0 = reinterpret_cast<float*>(smem + L::OFF_C0);
  float* c1 = reinterpret_cast<float*>(smem + L::OFF_C1);
  float* outb = reinterpret_cast<float*>(smem + L::OFF_OUTB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld0 = S + p.C + p.P;
  const int64_t T = p.T;

  // ---- scaling exponents: one for the recurrent matrices (their products share accumulators), one for W_out
  if (tid == 0) bars->amax[0] = bars->amax[1] = 0u;
  __syncthreads();
  {
    float mx = 0.f, mo = 0.f;
    for (int m = 0; m < NMAT; ++m) {
      const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
      for (int idx = tid; idx < 192 * 64; idx += kFwdThreads) mx = fmaxf(mx, fabsf(src[idx]));
    }
    for (int idx = tid; idx < NOUT * 64; idx += kFwdThreads) mo = fmaxf(mo, fabsf(p.out_w[idx]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    }
    if (lane == 0) {
      atomicMax(&bars->amax[0], __float_as_uint(mx));
      atomicMax(&bars->amax[1], __float_as_uint(mo));
    }
  }
  __syncthreads();
  const int ew = scale_exp(bars->amax[0]), eo = scale_exp(bars->amax[1]);
  const float w_scale = exp2i(ew), o_scale = exp2i(eo);
  const float sc = exp2i(-(ew + kHExp)), sco = exp2i(-(eo + kHExp));

  // ---- resident weight tiles (fp16 hi / lo, K-major, 128-byte swizzle)
  for (int m = 0; m < NMAT; ++m) {
    const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
    uint8_t* thi = smem + L::OFF_W + (m * 2) * kWTileBytes;
    uint8_t* tlo = thi + kWTileBytes;
    for (int idx = tid; idx < 192 * 8; idx += kFwdThreads) {
      const int n = idx >> 3, c = idx & 7;
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = src[n * 64 + c * 8 + q] * w_scale;
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(thi + sw128(n, c)) = hi;
      *reinterpret_cast<uint4*>(tlo + sw128(n, c)) = lo;
    }
  }
  for (int idx = tid; idx < 16 * 8; idx += kFwdThreads) {
    const int n = idx >> 3, c = idx & 7;
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = n < NOUT ? p.out_w[n * 64 + c * 8 + q] * o_scale : 0.f;
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(smem + L::OFF_WOUT + sw128(n, c)) = hi;
    *reinterpret_cast<uint4*>(smem + L::OFF_WOUT + kOutTileBytes + sw128(n, c)) = lo;
  }
  for (int idx = tid; idx < 64 * CS; idx += kFwdThreads) {
    const int j = idx / CS, q = idx % CS;
    float v = 0.f;
    if (q < 3 * S) v = p.w_ih[0][(int64_t)((q / S) * 64 + j) * ld0 + (q % S)];
    else if (q == 3 * S) v = p.b_hh[0][128 + j];
    c0[idx] = v;
  }
  if (NL == 2) {
    for (int j = tid; j < 64; j += kFwdThreads) {
      c1[j * 4 + 0] = p.b_ih[1][j] + p.b_hh[1][j];
      c1[j * 4 + 1] = p.b_ih[1][64 + j] + p.b_hh[1][64 + j];
      c1[j * 4 + 2] = p.b_ih[1][128 + j];
      c1[j * 4 + 3] = p.b_hh[1][128 + j];
    }
  }
  if (tid < 16) outb[tid] = tid < NOUT ? p.out_b[tid] : 0.f;

  if (tid == 0) {
    mbar_init(&bars->init, kEpiThreads);
    mbar_init(&bars->d0, 1);
    mbar_init(&bars->a0, kEpiThreads);
    mbar_init(&bars->d1, 1);
    mbar_init(&bars->a1, kEpiThreads);
    mbar_init(&bars->out, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  fence_proxy_async();  // the weight tiles were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;

  // ---- MMA issue (lane 0 of warp 0, between its epilogue phases) ---------------------------------
  const uint32_t w0 = smem_u32(smem + L::OFF_W);
  auto whi = [&](int m) { return w0 + (uint32_t)(m * 2) * kWTileBytes; };
  auto wlo = [&](int m) { return w0 + (uint32_t)(m * 2 + 1) * kWTileBytes; };
  const uint32_t a0h = smem_u32(smem + L::OFF_A), a0l = a0h + kATileBytes;
  const uint32_t a1h = a0h + 2 * kATileBytes, a1l = a1h + kATileBytes;
  const uint32_t woh = smem_u32(smem + L::OFF_WOUT), wol = woh + kOutTileBytes;
  constexpr uint32_t ID192 = idesc_f16(192), ID128 = idesc_f16(128), ID64 = idesc_f16(64), ID16 = idesc_f16(16);
  constexpr uint32_t NROWS = 128 * 128;  // byte offset of gate rows 128.. (the n block) in a weight tile
  auto issue_recurrent_l0 = [&]() {  // D0 = h0 . W_hh_l0^T (for the next step)
    issue_gemm(tmem + D0_COL, a0h, a0l, whi(0), wlo(0), ID192, false);
    umma_commit(&bars->d0);
  };
  auto issue_recurrent_l1 = [&]() {  // D1[r, u] = h1 . W_hh_l1[r, u]^T, D1[n_h] = h1 . W_hh_l1[n]^T
    issue_gemm(tmem + D1_COL, a1h, a1l, whi(2), wlo(2), ID128, false);
    issue_gemm(tmem + D1_COL + 192, a1h, a1l, whi(2) + NROWS, wlo(2) + NROWS, ID64, false);
  };
  auto issue_after_layer0 = [&](bool has_next) {
    if (NL == 2) {
      // layer 1 input part: r, u accumulate onto the recurrent part, n_i has its own columns
      issue_gemm(tmem + D1_COL, a0h, a0l, whi(1), wlo(1), ID128, true);
      issue_gemm(tmem + D1_COL + 128, a0h, a0l, whi(1) + NROWS, wlo(1) + NROWS, ID64, false);
      umma_commit(&bars->d1);
    } else {
      issue_gemm(tmem + DOUT_COL, a0h, a0l, woh, wol, ID16, false);
      umma_commit(&bars->out);
    }
    if (has_next) issue_recurrent_l0();
  };
  auto issue_after_layer1 = [&](bool has_next) {
    issue_gemm(tmem + DOUT_COL, a1h, a1l, woh, wol, ID16, false);
    umma_commit(&bars->out);
    if (has_next) issue_recurrent_l1();
  };
  uint32_t ph_init = 0, ph_a0 = 0, ph_a1 = 0;  // issuer-side phases (warp 0)

  {
    // ======================= gate epilogue =======================================================
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const int u0 = cg * kUPT;
    uint8_t* a_tiles = smem + L::OFF_A;
    const float hs = exp2i(kHExp);
    uint32_t ph_d0 = 0, ph_d1 = 0, ph_out = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t b_raw = tile * kTileRows + row;
      const bool ok = b_raw < p.B;
      const int64_t b = ok ? b_raw : p.B - 1;
      const bool writer = ok && cg == 0;
      // h(-1) = 0
#pragma unroll
      for (int k = 0; k < NL; ++k)
#pragma unroll
        for (int c = 0; c < kUPT / 8; ++c) {
          const uint32_t off = sw128(row, (u0 >> 3) + c);
          *reinterpret_cast<uint4*>(a_tiles + (2 * k) * kATileBytes + off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(a_tiles + (2 * k + 1) * kATileBytes + off) = make_uint4(0, 0, 0, 0);
        }
      fence_proxy_async();
      mbar_arrive(&bars->init);
      if (warp == 0) {
        mbar_wait(&bars->init, ph_init);
        ph_init ^= 1;
        tc_fence_after();
        if (lane == 0) {  // h(-1) = 0: the A tiles are zero, so these just clear the accumulators
          issue_recurrent_l0();
          if (NL == 2) issue_recurrent_l1();
        }
        __syncwarp();
      }

      float z[S], eps_cur[S], hprev[NL][kUPT];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        z[s] = p.x0[b * S + s];
        eps_cur[s] = p.eps[b * T * S + s];
        if (writer) p.paths[b * (T + 1) * S + s] = z[s];
      }
#pragma unroll
      for (int k = 0; k < NL; ++k)
#pragma unroll
        for (int j = 0; j < kUPT; ++j) hprev[k][j] = 0.f;
      // row-fastest tiled layouts (one coalesced 128-byte line per warp access):
      //   gi_ctx [tile][t][3H][128], stash [tile][t][NL][5][H][128]
      const float* gi_p = p.gi_ctx + tile * T * (int64_t)(192 * kTileRows) + (int64_t)u0 * kTileRows + row;
      float* st_p = p.stash ? p.stash + tile * T * (int64_t)(NL * kStashSlots * 64 * kTileRows) + (int64_t)u0 * kTileRows + row
                            : nullptr;
      const float* eps_p = p.eps + b * T * S;
      float* paths_o = p.paths + (b * (T + 1) + 1) * S;
      float* means_o = p.means + b * T * S;
      float* chol_o = p.chol + b * T * S * S;
      float* raw_o = p.raw ? p.raw + b * T * NTRIL : nullptr;

      // gi_ctx of this thread's 32 units, two register sets: units 0..15 of step t+1 are loaded while layer 1 of
      // step t runs, units 16..31 at the top of layer 0 (two 8-unit chunks before their first use)
      float g01[3][16], g23[3][16];
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int q = 0; q < 16; ++q) g01[g][q] = gi_p[(g * 64 + q) * kTileRows];

      for (int64_t t = 0; t < T; ++t) {
        const bool has_next = t + 1 < T;
        // ---------------- layer 0 ----------------
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int q = 0; q < 16; ++q) g23[g][q] = gi_p[(g * 64 + 16 + q) * kTileRows];
        mbar_wait(&bars->d0, ph_d0);
        ph_d0 ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kUPT / 8; ++c) {
          const int j0 = u0 + c * 8;
          uint32_t dr[8], du[8], dn[8];
          tmem_ld8_nowait(tl + D0_COL + j0, dr);
          tmem_ld8_nowait(tl + D0_COL + 64 + j0, du);
          tmem_ld8_nowait(tl + D0_COL + 128 + j0, dn);
          float gcur[3][8];
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int q = 0; q < 8; ++q) gcur[g][q] = c < 2 ? g01[g][(c & 1) * 8 + q] : g23[g][(c & 1) * 8 + q];
          tmem_ld_wait();
          float hx[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int jj = h4 * 4 + q, j = j0 + jj;
              float cc[CS];
#pragma unroll
              for (int v = 0; v < CS / 4; ++v) {
                const float4 w4 = *reinterpret_cast<const float4*>(c0 + j * CS + 4 * v);
                cc[4 * v] = w4.x; cc[4 * v + 1] = w4.y; cc[4 * v + 2] = w4.z; cc[4 * v + 3] = w4.w;
              }
              float pr = fmaf(sc, __uint_as_float(dr[jj]), gcur[0][jj]);
              float pu = fmaf(sc, __uint_as_float(du[jj]), gcur[1][jj]);
              float pni = gcur[2][jj];
              const float pnh = fmaf(sc, __uint_as_float(dn[jj]), cc[3 * S]);
#pragma unroll
              for (int s = 0; s < S; ++s) {
                pr = fmaf(cc[s], z[s], pr);
                pu = fmaf(cc[S + s], z[s], pu);
                pni = fmaf(cc[2 * S + s], z[s], pni);
              }
              const float r = sigmoid_f(pr);
              const float n = tanh_f(fmaf(r, pnh, pni));
              const float u = sigmoid_f(pu);
              const float hn = fmaf(u, hprev[0][c * 8 + jj] - n, n);
              hprev[0][c * 8 + jj] = hn;
              hx[jj] = hn * hs;
              if (st_p) {
                float* st = st_p + (c * 8 + jj) * kTileRows;
                st[kStashR * 64 * kTileRows] = r;
                st[kStashU * 64 * kTileRows] = u;
                st[kStashN * 64 * kTileRows] = n;
                st[kStashNhh * 64 * kTileRows] = pnh;
                st[kStashH * 64 * kTileRows] = hn;
              }
            }
          }
          uint4 hi, lo;
          split8(hx, hi, lo);
          const uint32_t off = sw128(row, j0 >> 3);
          *reinterpret_cast<uint4*>(a_tiles + off) = hi;
          *reinterpret_cast<uint4*>(a_tiles + kATileBytes + off) = lo;
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bars->a0);
        gi_p += 192 * kTileRows;
        if (warp == 0) {
          mbar_wait(&bars->a0, ph_a0);
          ph_a0 ^= 1;
          tc_fence_after();
          if (lane == 0) issue_after_layer0(has_next);
          __syncwarp();
        }
        if (has_next) {
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int q = 0; q < 16; ++q) g01[g][q] = gi_p[(g * 64 + q) * kTileRows];
        }

        // ---------------- layer 1 ----------------
        if (NL == 2) {
          mbar_wait(&bars->d1, ph_d1);
          ph_d1 ^= 1;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kUPT / 8; ++c) {
            const int j0 = u0 + c * 8;
            uint32_t dr[8], du[8], di[8], dn[8];
            tmem_ld8_nowait(tl + D1_COL + j0, dr);
            tmem_ld8_nowait(tl + D1_COL + 64 + j0, du);
            tmem_ld8_nowait(tl + D1_COL + 128 + j0, di);
            tmem_ld8_nowait(tl + D1_COL + 192 + j0, dn);
            tmem_ld_wait();
            float hx[8];
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int jj = h4 * 4 + q, j = j0 + jj;
                const float4 cb = *reinterpret_cast<const float4*>(c1 + j * 4);
                const float pr = fmaf(sc, __uint_as_float(dr[jj]), cb.x);
                const float pu = fmaf(sc, __uint_as_float(du[jj]), cb.y);
                const float pni = fmaf(sc, __uint_as_float(di[jj]), cb.z);
                const float pnh = fmaf(sc, __uint_as_float(dn[jj]), cb.w);
                const float r = sigmoid_f(pr);
                const float n = tanh_f(fmaf(r, pnh, pni));
                const float u = sigmoid_f(pu);
                const float hn = fmaf(u, hprev[NL - 1][c * 8 + jj] - n, n);
                hprev[NL - 1][c * 8 + jj] = hn;
                hx[jj] = hn * hs;
                if (st_p) {
                  float* st = st_p + (kStashSlots * 64 + c * 8 + jj) * kTileRows;
                  st[kStashR * 64 * kTileRows] = r;
                  st[kStashU * 64 * kTileRows] = u;
                  st[kStashN * 64 * kTileRows] = n;
                  st[kStashNhh * 64 * kTileRows] = pnh;
                  st[kStashH * 64 * kTileRows] = hn;
                }
              }
            }
            uint4 hi, lo;
            split8(hx, hi, lo);
            const uint32_t off = sw128(row, j0 >> 3);
            *reinterpret_cast<uint4*>(a_tiles + 2 * kATileBytes + off) = hi;
            *reinterpret_cast<uint4*>(a_tiles + 3 * kATileBytes + off) = lo;
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bars->a1);
          if (warp == 0) {
            mbar_wait(&bars->a1, ph_a1);
            ph_a1 ^= 1;
            tc_fence_after();
            if (lane == 0) issue_after_layer1(has_next);
            __syncwarp();
          }
        }
        if (st_p) st_p += NL * kStashSlots * 64 * kTileRows;

        // ---------------- output projection + reparameterised Euler-Maruyama update ----------------
        float eps_nxt[S];
        eps_p += S;
#pragma unroll
        for (int s = 0; s < S; ++s) eps_nxt[s] = has_next ? eps_p[s] : 0.f;
        mbar_wait(&bars->out, ph_out);
        ph_out ^= 1;
        tc_fence_after();
        float o[NOUT];
        {
          uint32_t ov[16];
          tmem_ld16_nowait(tl + DOUT_COL, ov);
          tmem_ld_wait();
#pragma unroll
          for (int m = 0; m < NOUT; ++m) o[m] = fmaf(sco, __uint_as_float(ov[m]), outb[m]);
        }
        tc_fence_before();
        float zn[S], Lm[NTRIL];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j <= s; ++j) {
            const int ti = s * (s + 1) / 2 + j;
            const float raw = o[S + ti];
            const float Lv = (j == s) ? fmaxf(raw, VISDE_DIAG_MIN) : raw;
            Lm[ti] = Lv;
            acc = fmaf(Lv, eps_cur[j], acc);
          }
          zn[s] = z[s] + o[s] * p.dt + acc * p.sqrt_dt;
        }
        if (writer) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            paths_o[s] = zn[s];
            means_o[s] = o[s];
#pragma unroll
            for (int j = 0; j < S; ++j) chol_o[s * S + j] = j <= s ? Lm[s * (s + 1) / 2 + j] : 0.f;
          }
          if (raw_o) {
#pragma unroll
            for (int ti = 0; ti < NTRIL; ++ti) raw_o[ti] = o[S + ti];
          }
        }
        paths_o += S;
        means_o += S;
        chol_o += S * S;
        if (raw_o) raw_o += NTRIL;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          z[s] = zn[s];
          eps_cur[s] = eps_nxt[s];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// gth[b, n] = b_ih_l0[n] + (n < 2H ? b_hh_l0[n] : 0) + sum_p theta[b, p] W_ih_l0[n, S + C + p]: the part of the
// layer-0 pre-activations that is constant along a trajectory; K0 adds it to gi_ctx as a per-row bias.
// Row-fastest tiled layout [ceil(B/128)][3H][128]; pad rows are zero.
__global__ void gth_kernel(const float* __restrict__ theta, const float* __restrict__ w_ih0, const float* __restrict__ b_ih0,
                           const float* __restrict__ b_hh0, int64_t B, int S, int C, int P, int H, float* __restrict__ gth) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int G = 3 * H;
  const int64_t ntile = (B + kTileRows - 1) / kTileRows;
  if (idx >= ntile * G * kTileRows) return;
  const int r = (int)(idx % kTileRows);
  const int n = (int)((idx / kTileRows) % G);
  const int64_t b = (idx / ((int64_t)kTileRows * G)) * kTileRows + r;
  float v = 0.f;
  if (b < B) {
    v = b_ih0[n] + (n < 2 * H ? b_hh0[n] : 0.f);
    const float* w = w_ih0 + (int64_t)n * (S + C + P) + S + C;
    for (int q = 0; q < P; ++q) v = fmaf(theta[b * P + q], w[q], v);
  }
  gth[idx] = v;
}

// out[b][t][f] = in[tile][t][f][row]: row-fastest tiled -> per-trajectory rows (bridge to the kernels that still
// read the [B, T, F] layouts)
__global__ void untile_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t B, int64_t T, int F) {
  __shared__ float tile[32][kTileRows + 1];
  const int f0 = blockIdx.x * 32;
  const int64_t t = blockIdx.y, tb = blockIdx.z;
  const float* src = in + ((tb * T + t) * F + f0) * (int64_t)kTileRows;
  for (int idx = threadIdx.x; idx < 32 * kTileRows; idx += blockDim.x) tile[idx / kTileRows][idx % kTileRows] = src[idx];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * kTileRows; idx += blockDim.x) {
    const int r = idx / 32, f = idx % 32;
    const int64_t b = tb * kTileRows + r;
    if (b < B && f0 + f < F) out[(b * T + t) * F + f0 + f] = tile[f][r];
  }
}

template <int NL, int S>
int launch_fwd_tc(const PathParams& p, cudaStream_t st) {
  const size_t smem = TcFwdSmem<NL, S>::bytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_tc_kernel<NL, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;
  path_fwd_tc_kernel<NL, S><<<(unsigned)(ntiles < sms ? ntiles : sms), kFwdThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

template <int NL>
int dispatch_s_tc(const PathParams& p, cudaStream_t st) {
  switch (p.S) {
    case 1: return launch_fwd_tc<NL, 1>(p, st);
    case 2: return launch_fwd_tc<NL, 2>(p, st);
    case 3: return launch_fwd_tc<NL, 3>(p, st);
    case 4: return launch_fwd_tc<NL, 4>(p, st);
  }
  set_error("tensor-core recurrence: unsupported state dim %d", p.S);
  return VISDE_EINVAL;
}

}  // namespace

bool tc_rec_supported(const PathParams& p) {
  return p.H == 64 && p.NL >= 1 && p.NL <= 2 && p.S >= 1 && p.S <= 4 && p.T >= 1 && p.B >= 1 &&
         p.T * (int64_t)(p.NL * kStashSlots * p.H) < (int64_t(1) << 31);
}

int launch_gth(const PathParams& p, float* gth, cudaStream_t st) {
  const int64_t n = ((p.B + kTileRows - 1) / kTileRows) * kTileRows * 3 * p.H;
  gth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.theta, p.w_ih[0], p.b_ih[0], p.b_hh[0], p.B, p.S, p.C, p.P, p.H, gth);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_untile(const float* in, float* out, int64_t B, int64_t T, int F, cudaStream_t st) {
  if (F % 32 != 0 || T > 65535) {
    set_error("untile: unsupported shape (F=%d, T=%lld)", F, (long long)T);
    return VISDE_EINVAL;
  }
  const int64_t ntile = (B + kTileRows - 1) / kTileRows;
  untile_kernel<<<dim3(F / 32, (unsigned)T, (unsigned)ntile), 256, 0, st>>>(in, out, B, T, F);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_path_fwd_tc(const PathParams& p, cudaStream_t st) {
  if (p.NL == 1) return dispatch_s_tc<1>(p, st);
  if (p.NL == 2) return dispatch_s_tc<2>(p, st);
  set_error("tensor-core recurrence: unsupported num_layers %d", p.NL);
  return VISDE_EINVAL;
}

}  // namespace visde
