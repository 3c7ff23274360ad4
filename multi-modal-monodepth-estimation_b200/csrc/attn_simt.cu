// Attention core, fp32 CUDA-core implementation (impl = 0): reference-precision path for fp32 parity,
// any window size up to 32x32, fp32 or bf16 storage.  The bf16 production path is attn_tc.cu (tcgen05).
//
// Replaces models/swin_transformer_v2.py:292-328 (cosine attention, bias, mask, softmax, P@V) with the
// block's pad/roll/partition/reverse/crop (:429-463) and the shift mask (:874-892) folded into the
// addressing: tokens are read from / written to the NATURAL [B,H,W,*] layout, nothing is permuted in HBM.
//
// One CTA per (window, head, row tile); one thread per row; the other side is streamed through shared
// memory in chunks of KC tokens with an online (flash-style) softmax, so S and P never exist in memory.
#include "common.cuh"
#include "wingeom.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

constexpr int HD = 32;          // head_dim of every Swin-V2 variant (models/model.py:18-29)
constexpr int KC = 64;          // streamed tokens per shared-memory chunk
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;   // the reference's -100 mask, in log2 units

struct AttnDims {
  WinGeom g;
  int C, nH, ntiles, nWm;
};

// per-window token table: tok[t] = flat natural token index or -1 (pad); meta[t] = koff | region << 16
__device__ __forceinline__ void build_token_table(const WinGeom& g, int64_t win, int* tok, int* meta) {
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1;
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    int b, i, j, si, sj;
    bool real = win_token(g, win, t, b, i, j, si, sj);
    tok[t] = real ? ((b * g.H + i) * g.W + j) : -1;
    int region = g.shift > 0 ? 3 * region_1d(si, g.Hp, ws, g.shift) + region_1d(sj, g.Wp, ws, g.shift) : 0;
    meta[t] = ((t / ws) * tw + (t % ws)) | (region << 16);
  }
}

template <typename T>
__device__ __forceinline__ void load_row32(const T* p, float (&v)[HD]) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    float t[4];
    ld4(p + c, t);
    v[c] = t[0]; v[c + 1] = t[1]; v[c + 2] = t[2]; v[c + 3] = t[3];
  }
}
template <typename T>
__device__ __forceinline__ void store_row32(T* p, const float (&v)[HD]) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    float t[4] = {v[c], v[c + 1], v[c + 2], v[c + 3]};
    st4(p + c, t);
  }
}
__device__ __forceinline__ float dot32(const float (&a)[HD], const float* __restrict__ b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    float4 t = *reinterpret_cast<const float4*>(b + c);
    s0 = fmaf(a[c], t.x, s0); s1 = fmaf(a[c + 1], t.y, s1);
    s2 = fmaf(a[c + 2], t.z, s2); s3 = fmaf(a[c + 3], t.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// cooperative load of two [kc][HD] fp32 tiles (8 threads x float4 per row) from natural-layout tensors.
// which: column offsets of the two parts; pad rows take padA / padB (nullable -> zeros).
template <typename T>
__device__ __forceinline__ void load_chunk(float* sa, float* sb, const T* __restrict__ srcA,
                                           const T* __restrict__ srcB, int64_t strideA, int64_t strideB,
                                           int colA, int colB, const float* __restrict__ padA,
                                           const float* __restrict__ padB, int padcol, const int* tok, int j0,
                                           int kc) {
  for (int idx = threadIdx.x; idx < kc * (HD / 4); idx += blockDim.x) {
    int r = idx >> 3, c4 = (idx & 7) * 4;
    int tj = tok[j0 + r];
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (tj >= 0) {
      ld4(srcA + (int64_t)tj * strideA + colA + c4, a);
      ld4(srcB + (int64_t)tj * strideB + colB + c4, b);
    } else {
      if (padA) ld4(padA + padcol + c4, a);
      if (padB) ld4(padB + padcol + c4, b);
    }
    *reinterpret_cast<float4*>(sa + r * HD + c4) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(sb + r * HD + c4) = make_float4(b[0], b[1], b[2], b[3]);
  }
}

// ------------------------------------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(256)
attn_fwd_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse,
                     const float* __restrict__ table16, const float* __restrict__ scale,
                     const float* __restrict__ qpad, const float* __restrict__ vpad,
                     const float* __restrict__ mask, AttnDims d) {
  const WinGeom& g = d.g;
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1, ntab = tw * tw, C = d.C, C3 = 3 * d.C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ks = reinterpret_cast<float*>(smem_raw);
  float* vs = ks + KC * HD;
  float* tab = vs + KC * HD;
  int* tok = reinterpret_cast<int*>(tab + ntab);
  int* meta = tok + N;

  const int64_t win = blockIdx.x / d.ntiles;
  const int tile = blockIdx.x - (int)(win * d.ntiles);
  const int h = blockIdx.y;
  build_token_table(g, win, tok, meta);
  for (int r = threadIdx.x; r < ntab; r += blockDim.x) tab[r] = table16[r * d.nH + h] * kLog2e;
  __syncthreads();

  const int i = tile * blockDim.x + threadIdx.x;
  const bool active = i < N;
  const float scale2 = scale[h] * kLog2e;
  int ti = -1, base_i = 0, reg_i = 0;
  float q[HD], o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) { q[c] = 0.f; o[c] = 0.f; }
  if (active) {
    ti = tok[i];
    base_i = (meta[i] & 0xffff) + (ws - 1) * (tw + 1);
    reg_i = meta[i] >> 16;
    if (ti >= 0) load_row32(qkv + (int64_t)ti * C3 + h * HD, q);
    else if (qpad) load_row32(qpad + h * HD, q);
  }
  const float* mrow = (mask && active) ? mask + ((win % d.nWm) * N + i) * (int64_t)N : nullptr;
  float m = -INFINITY, l = 0.f;

  for (int j0 = 0; j0 < N; j0 += KC) {
    const int kc = min(KC, N - j0);
    if (j0 > 0) __syncthreads();
    load_chunk<T>(ks, vs, qkv, qkv, C3, C3, C + h * HD, 2 * C + h * HD, nullptr, vpad, h * HD, tok, j0, kc);
    __syncthreads();
    if (active) {
      for (int jj = 0; jj < kc; jj += 8) {
        float s[8];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int j = jj + u;
          if (j < kc) {
            int mj = meta[j0 + j];
            float v = dot32(q, ks + j * HD) * scale2 + tab[base_i - (mj & 0xffff)];
            if ((mj >> 16) != reg_i) v += kMaskLog2;
            if (mrow) v += mrow[j0 + j] * kLog2e;
            s[u] = v;
            mx = fmaxf(mx, v);
          } else {
            s[u] = -INFINITY;
          }
        }
        const float m_new = fmaxf(m, mx);
        const float corr = exp2f(m - m_new);
        l *= corr;
#pragma unroll
        for (int c = 0; c < HD; ++c) o[c] *= corr;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int j = jj + u;
          if (j < kc) {
            float p = exp2f(s[u] - m_new);
            l += p;
            const float* vr = vs + j * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 4) {
              float4 t = *reinterpret_cast<const float4*>(vr + c);
              o[c] = fmaf(p, t.x, o[c]); o[c + 1] = fmaf(p, t.y, o[c + 1]);
              o[c + 2] = fmaf(p, t.z, o[c + 2]); o[c + 3] = fmaf(p, t.w, o[c + 3]);
            }
          }
        }
        m = m_new;
      }
    }
  }
  if (active) {
    lse[(win * d.nH + h) * N + i] = (m + log2f(l)) * kLn2;
    if (ti >= 0) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int c = 0; c < HD; ++c) o[c] *= inv;
      store_row32(out + (int64_t)ti * C + h * HD, o);
    }
  }
}

// ------------------------------------------------------------------------------- backward, query side
// thread = query row i; streams K,V.  Produces dq (through the F.normalize backward), dtable16, dscale.
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_dq_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                        const float* __restrict__ lse, const float* __restrict__ inv_norm,
                        const float* __restrict__ table16, const float* __restrict__ scale,
                        const float* __restrict__ qpad, const float* __restrict__ vpad,
                        const float* __restrict__ mask, T* __restrict__ dqkv, float* __restrict__ dtable16,
                        float* __restrict__ dscale, AttnDims d) {
  const WinGeom& g = d.g;
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1, ntab = tw * tw, C = d.C, C3 = 3 * d.C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ks = reinterpret_cast<float*>(smem_raw);
  float* vs = ks + KC * HD;
  float* tab = vs + KC * HD;
  float* dtab = tab + ntab;
  int* tok = reinterpret_cast<int*>(dtab + ntab);
  int* meta = tok + N;
  __shared__ float red[8];

  const int64_t win = blockIdx.x / d.ntiles;
  const int tile = blockIdx.x - (int)(win * d.ntiles);
  const int h = blockIdx.y;
  build_token_table(g, win, tok, meta);
  for (int r = threadIdx.x; r < ntab; r += blockDim.x) { tab[r] = table16[r * d.nH + h] * kLog2e; dtab[r] = 0.f; }
  __syncthreads();

  const int i = tile * blockDim.x + threadIdx.x;
  const float sc = scale[h], scale2 = sc * kLog2e;
  int ti = -1, base_i = 0, reg_i = 0;
  float q[HD], go[HD], dq[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) { q[c] = 0.f; go[c] = 0.f; dq[c] = 0.f; }
  float D = 0.f, lse2 = 0.f, dsc = 0.f;
  if (i < N) ti = tok[i];
  const bool active = ti >= 0;            // pad query rows are cropped: zero upstream gradient
  if (active) {
    base_i = (meta[i] & 0xffff) + (ws - 1) * (tw + 1);
    reg_i = meta[i] >> 16;
    load_row32(qkv + (int64_t)ti * C3 + h * HD, q);
    load_row32(dout + (int64_t)ti * C + h * HD, go);
    float ov[HD];
    load_row32(out + (int64_t)ti * C + h * HD, ov);
#pragma unroll
    for (int c = 0; c < HD; ++c) D = fmaf(go[c], ov[c], D);
    lse2 = lse[(win * d.nH + h) * N + i] * kLog2e;
  }
  const float* mrow = (mask && active) ? mask + ((win % d.nWm) * N + i) * (int64_t)N : nullptr;

  for (int j0 = 0; j0 < N; j0 += KC) {
    const int kc = min(KC, N - j0);
    if (j0 > 0) __syncthreads();
    load_chunk<T>(ks, vs, qkv, qkv, C3, C3, C + h * HD, 2 * C + h * HD, nullptr, vpad, h * HD, tok, j0, kc);
    __syncthreads();
    if (active) {
      for (int j = 0; j < kc; ++j) {
        const int mj = meta[j0 + j];
        const int rel = base_i - (mj & 0xffff);
        const float cosv = dot32(q, ks + j * HD);
        float s2 = cosv * scale2 + tab[rel];
        if ((mj >> 16) != reg_i) s2 += kMaskLog2;
        if (mrow) s2 += mrow[j0 + j] * kLog2e;
        const float p = exp2f(s2 - lse2);
        const float dp = dot32(go, vs + j * HD);
        const float ds = p * (dp - D);
        const float* kr = ks + j * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          float4 t = *reinterpret_cast<const float4*>(kr + c);
          dq[c] = fmaf(ds, t.x, dq[c]); dq[c + 1] = fmaf(ds, t.y, dq[c + 1]);
          dq[c + 2] = fmaf(ds, t.z, dq[c + 2]); dq[c + 3] = fmaf(ds, t.w, dq[c + 3]);
        }
        dsc = fmaf(ds, cosv, dsc);
        atomicAdd(dtab + rel, ds);
      }
    }
  }
  if (active) {
    // dq_hat = scale * dS k_hat;  dq = (dq_hat - q_hat <dq_hat, q_hat>) / max(|q|, eps)
    float dotq = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) { dq[c] *= sc; dotq = fmaf(dq[c], q[c], dotq); }
    // no inv_norm tensor: q / k were not normalised (attn_type='normal'), the gradient passes through
    const float invn = inv_norm ? inv_norm[((int64_t)ti * 2 + 0) * d.nH + h] : 1.f;
    if (!inv_norm) dotq = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] = (dq[c] - q[c] * dotq) * invn;
    store_row32(dqkv + (int64_t)ti * C3 + h * HD, dq);
  }
  dsc = warp_sum(dsc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dsc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += red[w];
    atomicAdd(dscale + h, s);
  }
  for (int r = threadIdx.x; r < ntab; r += blockDim.x) {
    float v = dtab[r];
    if (v != 0.f) atomicAdd(dtable16 + r * d.nH + h, v);
  }
}

// --------------------------------------------------------------------------------- backward, key side
// thread = key row j; streams Q, dO (+ lse, D).  Produces dk (through the normalize backward) and dv.
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_dkv_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                         const float* __restrict__ lse, const float* __restrict__ inv_norm,
                         const float* __restrict__ table16, const float* __restrict__ scale,
                         const float* __restrict__ vpad, const float* __restrict__ mask, T* __restrict__ dqkv,
                         float* __restrict__ dvpad, AttnDims d) {
  const WinGeom& g = d.g;
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1, ntab = tw * tw, C = d.C, C3 = 3 * d.C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);
  float* gos = qs + KC * HD;
  float* tab = gos + KC * HD;
  float* lses = tab + ntab;        // [KC] lse in log2 units (+inf for pad query rows -> p = 0)
  float* Ds = lses + KC;           // [KC]
  int* tok = reinterpret_cast<int*>(Ds + KC);
  int* meta = tok + N;

  const int64_t win = blockIdx.x / d.ntiles;
  const int tile = blockIdx.x - (int)(win * d.ntiles);
  const int h = blockIdx.y;
  build_token_table(g, win, tok, meta);
  for (int r = threadIdx.x; r < ntab; r += blockDim.x) tab[r] = table16[r * d.nH + h] * kLog2e;
  __syncthreads();

  const int j = tile * blockDim.x + threadIdx.x;
  const bool active = j < N;
  const float sc = scale[h], scale2 = sc * kLog2e;
  int tj = -1, koff_j = 0, reg_j = 0;
  float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) { k[c] = 0.f; v[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
  if (active) {
    tj = tok[j];
    koff_j = meta[j] & 0xffff;
    reg_j = meta[j] >> 16;
    if (tj >= 0) {
      load_row32(qkv + (int64_t)tj * C3 + C + h * HD, k);
      load_row32(qkv + (int64_t)tj * C3 + 2 * C + h * HD, v);
    } else if (vpad) {
      load_row32(vpad + h * HD, v);
    }
  }
  const float* mcol = (mask && active) ? mask + (win % d.nWm) * (int64_t)N * N + j : nullptr;
  const int base_add = (ws - 1) * (tw + 1);

  for (int i0 = 0; i0 < N; i0 += KC) {
    const int kc = min(KC, N - i0);
    if (i0 > 0) __syncthreads();
    // load q_hat and dO rows of the chunk; D_i = <dO_i, O_i> reduced over the 8 threads of a row
    for (int idx = threadIdx.x; idx < ((kc * 8 + 31) & ~31); idx += blockDim.x) {
      int r = idx >> 3, c4 = (idx & 7) * 4;
      int ti = (r < kc) ? tok[i0 + r] : -1;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, o[4] = {0.f, 0.f, 0.f, 0.f};
      if (ti >= 0) {
        ld4(qkv + (int64_t)ti * C3 + h * HD + c4, a);
        ld4(dout + (int64_t)ti * C + h * HD + c4, b);
        ld4(out + (int64_t)ti * C + h * HD + c4, o);
      }
      float part = b[0] * o[0] + b[1] * o[1] + b[2] * o[2] + b[3] * o[3];
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (r < kc) {
        *reinterpret_cast<float4*>(qs + r * HD + c4) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(gos + r * HD + c4) = make_float4(b[0], b[1], b[2], b[3]);
        if ((idx & 7) == 0) {
          Ds[r] = part;
          lses[r] = (ti >= 0) ? lse[(win * d.nH + h) * N + i0 + r] * kLog2e : INFINITY;
        }
      }
    }
    __syncthreads();
    if (active) {
      for (int r = 0; r < kc; ++r) {
        const int mi = meta[i0 + r];
        const float* qr = qs + r * HD;
        const float* gr = gos + r * HD;
        float s2 = dot32(k, qr) * scale2 + tab[(mi & 0xffff) + base_add - koff_j];
        if ((mi >> 16) != reg_j) s2 += kMaskLog2;
        if (mcol) s2 += mcol[(int64_t)(i0 + r) * N] * kLog2e;
        const float p = exp2f(s2 - lses[r]);
        const float dp = dot32(v, gr);
        const float ds = p * (dp - Ds[r]);
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          float4 tq = *reinterpret_cast<const float4*>(qr + c);
          float4 tg = *reinterpret_cast<const float4*>(gr + c);
          dk[c] = fmaf(ds, tq.x, dk[c]); dk[c + 1] = fmaf(ds, tq.y, dk[c + 1]);
          dk[c + 2] = fmaf(ds, tq.z, dk[c + 2]); dk[c + 3] = fmaf(ds, tq.w, dk[c + 3]);
          dv[c] = fmaf(p, tg.x, dv[c]); dv[c + 1] = fmaf(p, tg.y, dv[c + 1]);
          dv[c + 2] = fmaf(p, tg.z, dv[c + 2]); dv[c + 3] = fmaf(p, tg.w, dv[c + 3]);
        }
      }
    }
  }
  if (active) {
    if (tj >= 0) {
      float dotk = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) { dk[c] *= sc; dotk = fmaf(dk[c], k[c], dotk); }
      const float invn = inv_norm ? inv_norm[((int64_t)tj * 2 + 1) * d.nH + h] : 1.f;
      if (!inv_norm) dotk = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) dk[c] = (dk[c] - k[c] * dotk) * invn;
      store_row32(dqkv + (int64_t)tj * C3 + C + h * HD, dk);
      store_row32(dqkv + (int64_t)tj * C3 + 2 * C + h * HD, dv);
    } else if (dvpad) {
      // pad token: v = v_bias, so its dV reaches v_bias (the reference gets this through F.pad + bias add)
#pragma unroll
      for (int c = 0; c < HD; ++c) atomicAdd(dvpad + h * HD + c, dv[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------- host side
static int attn_dims(AttnDims* d, int* threads, int B, int H, int W, int C, int nH, int ws, int shift, int nWm,
                     const void* mask) {
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && nH > 0 && ws > 0, "attn: non-positive dimension");
  BSW_REQUIRE(C == nH * HD, "attn: head_dim must be 32 (C=%d, nH=%d)", C, nH);
  BSW_REQUIRE(ws <= 32, "attn: window size %d > 32 unsupported", ws);
  BSW_REQUIRE(shift >= 0 && shift < ws, "attn: shift %d must be in [0, ws)", shift);
  BSW_REQUIRE(!mask || nWm > 0, "attn: explicit mask needs nWm > 0");
  BSW_REQUIRE((int64_t)B * H * W < (1ll << 31) / 4, "attn: too many tokens for 32-bit token indices");
  make_geom(&d->g, B, H, W, ws, shift);
  d->C = C; d->nH = nH; d->nWm = nWm > 0 ? nWm : 1;
  int N = ws * ws;
  d->ntiles = (N + 255) / 256;
  int per = (N + d->ntiles - 1) / d->ntiles;
  *threads = (per + 31) / 32 * 32;
  return B200SWIN_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) BSW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B200SWIN_OK;
}

template <typename T>
static int attn_fwd_simt(const void* qkv, void* out, float* lse, const float* table16, const float* scale,
                         const float* qpad, const float* vpad, const float* mask, const AttnDims& d, int threads,
                         cudaStream_t st) {
  const int N = d.g.ws * d.g.ws, ntab = (2 * d.g.ws - 1) * (2 * d.g.ws - 1);
  size_t smem = (size_t)(2 * KC * HD + ntab) * 4 + (size_t)2 * N * 4;
  int rc = set_smem(attn_fwd_simt_kernel<T>, smem);
  if (rc) return rc;
  int64_t nwin = (int64_t)d.g.B * d.g.nWh * d.g.nWw;
  dim3 grid((unsigned)(nwin * d.ntiles), d.nH);
  attn_fwd_simt_kernel<T><<<grid, threads, smem, st>>>((const T*)qkv, (T*)out, lse, table16, scale, qpad, vpad, mask, d);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

template <typename T>
static int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, const float* inv_norm,
                         const float* table16, const float* scale, const float* qpad, const float* vpad,
                         const float* mask, void* dqkv, float* dtable16, float* dscale, float* dvpad,
                         const AttnDims& d, int threads, cudaStream_t st) {
  const int N = d.g.ws * d.g.ws, ntab = (2 * d.g.ws - 1) * (2 * d.g.ws - 1);
  int64_t nwin = (int64_t)d.g.B * d.g.nWh * d.g.nWw;
  dim3 grid((unsigned)(nwin * d.ntiles), d.nH);
  size_t smem_q = (size_t)(2 * KC * HD + 2 * ntab) * 4 + (size_t)2 * N * 4;
  int rc = set_smem(attn_bwd_dq_simt_kernel<T>, smem_q);
  if (rc) return rc;
  attn_bwd_dq_simt_kernel<T><<<grid, threads, smem_q, st>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, inv_norm,
                                                         table16, scale, qpad, vpad, mask, (T*)dqkv, dtable16, dscale, d);
  BSW_LAUNCH_CHECK();
  size_t smem_k = (size_t)(2 * KC * HD + ntab + 2 * KC) * 4 + (size_t)2 * N * 4;
  rc = set_smem(attn_bwd_dkv_simt_kernel<T>, smem_k);
  if (rc) return rc;
  attn_bwd_dkv_simt_kernel<T><<<grid, threads, smem_k, st>>>((const T*)qkv, (const T*)out, (const T*)dout, lse,
                                                          inv_norm, table16, scale, vpad, mask, (T*)dqkv, dvpad, d);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

// implemented in attn_tc.cu (tcgen05 path); returns B200SWIN_EINVAL with a message when the shape is unsupported
int attn_fwd_tc(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, const float* mask, int nWm, int B, int H, int W, int C, int nH, int ws, int shift,
                int family, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                const float* table16, const float* scale, const float* qpad, const float* vpad, const float* mask,
                int nWm, void* dqkv, float* dtable16, float* dscale, float* dvpad, float* dcol, void* workspace,
                size_t workspace_bytes, int B, int H, int W, int C, int nH, int ws, int shift, int family,
                cudaStream_t st);
bool attn_bwd_tc_colsum_supported(int ws, int family);
size_t attn_bwd_tc_workspace_bytes(int B, int H, int W, int nH, int ws);

}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_attn_fwd(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale,
                                 const float* qpad, const float* vpad, const float* mask, int nWm, int B, int H, int W,
                                 int C, int nH, int ws, int shift, int dtype, int impl, void* stream) {
  BSW_REQUIRE(qkv && out && lse && table16 && scale, "attn_fwd: null pointer");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "attn_fwd: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (impl >= 1 && impl <= 4) {
    BSW_REQUIRE(dtype == B200SWIN_BF16, "attn_fwd: the tensor-core path stores bf16");
    return attn_fwd_tc(qkv, out, out_lo, lse, table16, scale, qpad, vpad, mask, nWm, B, H, W, C, nH, ws, shift, impl - 1, st);
  }
  BSW_REQUIRE(impl == 0, "attn_fwd: unknown impl %d", impl);
  (void)out_lo;                                     // the CUDA-core path is the reference-precision path: fp32 tensors, nothing to split
  AttnDims d;
  int threads;
  int rc = attn_dims(&d, &threads, B, H, W, C, nH, ws, shift, nWm, mask);
  if (rc) return rc;
  if (dtype == B200SWIN_F32) return attn_fwd_simt<float>(qkv, out, lse, table16, scale, qpad, vpad, mask, d, threads, st);
  return attn_fwd_simt<__nv_bfloat16>(qkv, out, lse, table16, scale, qpad, vpad, mask, d, threads, st);
}

extern "C" size_t b200swin_attn_bwd_workspace_bytes(int B, int H, int W, int nH, int ws, int dtype, int impl) {
  // scratch of the backward: D = <dO, O> per (token, head) for the warp-specialised tensor-core kernel
  if (impl < 1 || impl > 4 || dtype != B200SWIN_BF16 || B <= 0 || H <= 0 || W <= 0 || nH <= 0) return 0;
  return attn_bwd_tc_workspace_bytes(B, H, W, nH, ws);
}

extern "C" int b200swin_attn_bwd_colsum_supported(int ws, int dtype, int impl) {
  return (dtype == B200SWIN_BF16 && impl >= 1 && impl <= 4 && attn_bwd_tc_colsum_supported(ws, impl - 1)) ? 1 : 0;
}

extern "C" int b200swin_attn_bwd(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse,
                                 const float* inv_norm, const float* table16, const float* scale, const float* qpad,
                                 const float* vpad, const float* mask, int nWm, void* dqkv, float* dtable16,
                                 float* dscale, float* dvpad, float* dqkv_colsum, int B, int H, int W, int C, int nH,
                                 int ws, int shift, int dtype, int impl, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  BSW_REQUIRE(qkv && out && dout && lse && table16 && scale && dqkv && dtable16 && dscale, "attn_bwd: null pointer");
  BSW_REQUIRE(inv_norm || impl == 0 || impl == 2,
              "attn_bwd: inv_norm = NULL (un-normalised q, k: attn_type='normal') is served by impl 0 and 2 only");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "attn_bwd: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (impl >= 1 && impl <= 4) {
    BSW_REQUIRE(dtype == B200SWIN_BF16, "attn_bwd: the tensor-core path stores bf16");
    return attn_bwd_tc(qkv, out, out_lo, dout, lse, inv_norm, table16, scale, qpad, vpad, mask, nWm, dqkv, dtable16, dscale,
                       dvpad, dqkv_colsum, workspace, workspace_bytes, B, H, W, C, nH, ws, shift, impl - 1, st);
  }
  BSW_REQUIRE(impl == 0, "attn_bwd: unknown impl %d", impl);
  BSW_REQUIRE(!dqkv_colsum, "attn_bwd: the CUDA-core kernels do not produce the column sums");
  (void)out_lo;
  AttnDims d;
  int threads;
  int rc = attn_dims(&d, &threads, B, H, W, C, nH, ws, shift, nWm, mask);
  if (rc) return rc;
  if (dtype == B200SWIN_F32)
    return attn_bwd_simt<float>(qkv, out, dout, lse, inv_norm, table16, scale, qpad, vpad, mask, dqkv, dtable16, dscale,
                                dvpad, d, threads, st);
  return attn_bwd_simt<__nv_bfloat16>(qkv, out, dout, lse, inv_norm, table16, scale, qpad, vpad, mask, dqkv, dtable16,
                                      dscale, dvpad, d, threads, st);
}
