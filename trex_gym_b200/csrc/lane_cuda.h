// lane_cuda.h -- native sm_100a backend of the lane-vector vocabulary used by trex_core.h.
//
// trex_core.h is written once against a tiny SPMD vocabulary (vf/vi/vb = "one value per
// lane of the warp that owns this environment").  Here, on the device, a vf is simply the
// thread's own float register and the collective operations are warp shuffles / votes.
// tests/emu/lane_emu.h provides the same vocabulary as 32-wide host structs so that the
// *identical* kernel source can be executed lane-for-lane on a CPU in the test-suite
// (there is no GPU in the build container).  The emulator is test infrastructure only: it
// is not compiled into libtrex_b200.so and there is no CPU fallback in the product.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TREX_FN __device__ __forceinline__
#define TREX_UNROLL _Pragma("unroll")
#define TREX_ROLLED _Pragma("unroll 1")
#define TREX_FULL 0xffffffffu

typedef float vf;
typedef int vi;
typedef bool vb;

TREX_FN vi lane_id() { return (int)(threadIdx.x & 31u); }
TREX_FN vf vbroadcast(float x) { return x; }
TREX_FN vf sel(vb p, vf a, vf b) { return p ? a : b; }
TREX_FN vi seli(vb p, vi a, vi b) { return p ? a : b; }
TREX_FN vf shfl(vf x, int src) { return __shfl_sync(TREX_FULL, x, src); }
TREX_FN vf shflv(vf x, vi src) { return __shfl_sync(TREX_FULL, x, src); }
TREX_FN vf shfl_xor(vf x, int m) { return __shfl_xor_sync(TREX_FULL, x, m); }
TREX_FN float lane_value(vf x, int lane) { return __shfl_sync(TREX_FULL, x, lane); }
TREX_FN int lane_value_i(vi x, int lane) { return __shfl_sync(TREX_FULL, x, lane); }
TREX_FN uint32_t vballot(vb p) { return __ballot_sync(TREX_FULL, p); }
TREX_FN bool vany(vb p) { return __any_sync(TREX_FULL, p); }
TREX_FN void warp_sync() { __syncwarp(); }

// butterfly reductions: every lane ends with the same bits (fixed association order)
TREX_FN vf warp_sum(vf x) {
  TREX_UNROLL for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(TREX_FULL, x, m);
  return x;
}
TREX_FN vf warp_max(vf x) {
  TREX_UNROLL for (int m = 16; m > 0; m >>= 1) x = fmaxf(x, __shfl_xor_sync(TREX_FULL, x, m));
  return x;
}

// memory (global or shared): per-lane index
TREX_FN vf ld(const float* p, vi idx) { return p[idx]; }
TREX_FN vf ldg_ro(const float* p, vi idx) { return __ldg(p + idx); }
TREX_FN vf ld_if(const float* p, vi idx, vb pred, float dflt) { return pred ? p[idx] : dflt; }
TREX_FN void st(float* p, vi idx, vf v) { p[idx] = v; }
TREX_FN void st_if(float* p, vi idx, vf v, vb pred) { if (pred) p[idx] = v; }
TREX_FN void st_u8_if(uint8_t* p, vi idx, vi v, vb pred) { if (pred) p[idx] = (uint8_t)v; }
// uniform (same address on every lane) loads
TREX_FN float ldu(const float* p, int idx) { return p[idx]; }
TREX_FN int ldui(const int* p, int idx) { return p[idx]; }

TREX_FN vf vfma(vf a, vf b, vf c) { return fmaf(a, b, c); }
TREX_FN vf vsqrt(vf x) { return sqrtf(x); }
TREX_FN vf vabs(vf x) { return fabsf(x); }
TREX_FN vf vmin(vf a, vf b) { return fminf(a, b); }
TREX_FN vf vmax(vf a, vf b) { return fmaxf(a, b); }
// symmetric clamp min(max(x, -m), m) for m >= 0 as ONE instruction: FMNMX.XORSIGN takes min(|x|, |m|) with the sign bit
// sign(x) ^ sign(m) (half the latency of the FMNMX pair on the Gauss-Seidel chain; same bits for every non-NaN x)
TREX_FN vf vclamp_sym(vf x, vf m) { float r; asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(m)); return r; }
// Packed FP32 pairs (Blackwell FFMA2 / FMUL2: `fma.rn.f32x2`, one issue slot for two IEEE fused multiply-adds on an aligned
// register pair -- the same bits as two FFMAs).  The solvers are issue bound; their impulse publications update two or four
// accumulators with one scalar.  (c0, c1) += (a0, a1) * s / (c0, c1) += (a0, a1) * (b0, b1) / (r0, r1) = (a0, a1) * s
TREX_FN void vfma2s(vf& c0, vf& c1, vf a0, vf a1, vf s) {
  asm("{\n .reg .b64 ra, rs, rc;\n mov.b64 ra, {%2, %3};\n mov.b64 rs, {%4, %4};\n mov.b64 rc, {%0, %1};\n"
      " fma.rn.f32x2 rc, ra, rs, rc;\n mov.b64 {%0, %1}, rc;\n}" : "+f"(c0), "+f"(c1) : "f"(a0), "f"(a1), "f"(s));
}
TREX_FN void vfma2v(vf& c0, vf& c1, vf a0, vf a1, vf b0, vf b1) {
  asm("{\n .reg .b64 ra, rb, rc;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%0, %1};\n"
      " fma.rn.f32x2 rc, ra, rb, rc;\n mov.b64 {%0, %1}, rc;\n}" : "+f"(c0), "+f"(c1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
TREX_FN void vmul2s(vf& r0, vf& r1, vf a0, vf a1, vf s) {
  asm("{\n .reg .b64 ra, rs, rr;\n mov.b64 ra, {%2, %3};\n mov.b64 rs, {%4, %4};\n"
      " mul.rn.f32x2 rr, ra, rs;\n mov.b64 {%0, %1}, rr;\n}" : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(s));
}
TREX_FN vf vsin(vf x) { return sinf(x); }
TREX_FN vf vcos(vf x) { return cosf(x); }
TREX_FN vf vdiv(vf a, vf b) { return a / b; }
TREX_FN vb visnan(vf x) { return !(fabsf(x) <= 3.0e38f); }  // NaN or inf
TREX_FN vi vf2i_bits(vf x) { return __float_as_int(x); }
TREX_FN vf vi2f_bits(vi x) { return __int_as_float(x); }
TREX_FN vf vi2f(vi x) { return (float)x; }

TREX_FN vi ldi(const int* p, vi idx) { return p[idx]; }
TREX_FN void sti_if(int* p, vi idx, vi v, vb pred) { if (pred) p[idx] = v; }
TREX_FN vi rank_below(uint32_t mask) { return __popc(mask & ((1u << (threadIdx.x & 31u)) - 1u)); }
TREX_FN int popc_u(uint32_t m) { return __popc(m); }
// non-contracted arithmetic (the reward must be reproducible bit-for-bit from the outputs)
TREX_FN float fmul_rn(float a, float b) { return __fmul_rn(a, b); }
TREX_FN float fadd_rn(float a, float b) { return __fadd_rn(a, b); }
TREX_FN vf vmul_rn(vf a, vf b) { return __fmul_rn(a, b); }
TREX_FN int ctz_u(uint32_t m) { return __ffs((int)m) - 1; }
TREX_FN int clz_u(uint32_t m) { return __clz((int)m); }
// rsqrt.approx.ftz: one MUFU.RSQ (rsqrtf() adds a denormal pre/post-scaling pair around it; the solvers only take the
// reciprocal root of a squared impulse norm already known to be > 0, where a denormal argument may flush)
TREX_FN vf vrsqrt(vf x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
TREX_FN void stb(unsigned char* p, vi idx, vi v) { p[idx] = (unsigned char)v; }
TREX_FN vi ldb(const unsigned char* p, vi idx) { return (int)p[idx]; }
TREX_FN long long cycle_count() { return clock64(); }
TREX_FN void cta_sync() { __syncthreads(); }  // every warp of the CTA, the same number of times

// Philox4x32-10 -> 4 uniforms in [0,1) per lane (counter-based: reset sampler)
TREX_FN void philox4_uniform(vi c0, vi c1, vi c2, vi c3, uint32_t k0, uint32_t k1, vf out[4]) {
  uint32_t c[4] = {(uint32_t)c0, (uint32_t)c1, (uint32_t)c2, (uint32_t)c3};
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = (float)(c[k] >> 8) * (1.0f / 16777216.0f);
}

TREX_FN vi vmini(vi a, int b) { return a < b ? a : b; }
TREX_FN vi vf2i(vf x) { return (int)x; }
TREX_FN vi warp_maxi(vi x) {
  TREX_UNROLL for (int m = 16; m > 0; m >>= 1) { const int y = __shfl_xor_sync(TREX_FULL, x, m); x = x > y ? x : y; }
  return x;
}
// width-8 lane groups (four environments per warp in solve4)
TREX_FN vf shfl_group8(vf x, int src) { return __shfl_sync(TREX_FULL, x, src, 8); }
TREX_FN vf shflv_group8(vf x, vi src) { return __shfl_sync(TREX_FULL, x, src, 8); }  // per-lane source (uniform within a group)
TREX_FN vf group8_sum(vf x) { TREX_UNROLL for (int m = 4; m > 0; m >>= 1) x += __shfl_xor_sync(TREX_FULL, x, m); return x; }
TREX_FN vf group8_max(vf x) { TREX_UNROLL for (int m = 4; m > 0; m >>= 1) x = fmaxf(x, __shfl_xor_sync(TREX_FULL, x, m)); return x; }

// per-lane integer helpers of the active-set signature (trex_core.h: StepStats)
TREX_FN vi shfl_xor_i(vi x, int m) { return __shfl_xor_sync(TREX_FULL, x, m); }
TREX_FN vi sig_mix_v(vi h, vi w) { return (int)(((uint32_t)h ^ (uint32_t)w) * 16777619u); }
// width-16 lane groups (two environments per warp in solve2)
TREX_FN vf shfl_group16(vf x, int src) { return __shfl_sync(TREX_FULL, x, src, 16); }
TREX_FN vf shflv_group16(vf x, vi src) { return __shfl_sync(TREX_FULL, x, src, 16); }
TREX_FN vf group16_sum(vf x) { TREX_UNROLL for (int m = 8; m > 0; m >>= 1) x += __shfl_xor_sync(TREX_FULL, x, m); return x; }
// 2 consecutive floats per lane (8-byte aligned offset): one 64-bit access
TREX_FN void ld2(const float* p, vi idx, vf out[2]) {
  const float2 t = *reinterpret_cast<const float2*>(p + idx);
  out[0] = t.x; out[1] = t.y;
}
TREX_FN void st2_if(float* p, vi idx, const vf v[2], vb pred) {
  if (pred) *reinterpret_cast<float2*>(p + idx) = make_float2(v[0], v[1]);
}

// Tensor memory (TMEM, 256 KB per SM on sm_100a) as a software-managed per-lane scratchpad: `tcgen05.st/ld ... 32x32b.x4`
// moves four consecutive 32-bit columns between registers and the TMEM lane of each thread of the warp (a warp reaches the
// 32 lanes 32 * (warp % 4) ..).  solve2 keeps the Delassus matrix of its two environments there: lane c' holds its own
// three entries of every row r in columns 4 r .. 4 r + 2.  Loads are asynchronous: tmem_wait4 passes the registers
// through `tcgen05.wait::ld`, so no use can be scheduled ahead of it.
struct tmem_t { uint32_t base; };  // TMEM address of column 0 of this warp's lane slice
TREX_FN void tmem_st4(tmem_t t, int col, vf a, vf b, vf c, vf d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t.base + (uint32_t)col), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// one warp of the CTA allocates `cols` columns (power of two >= 32) and publishes the address through shared memory; every thread
// of the CTA calls this (it contains a CTA barrier); the same warp frees them at the end (tmem_free_cta, after a CTA barrier)
template <int COLS>
__device__ __forceinline__ uint32_t tmem_alloc_cta(uint32_t* smem_slot) {
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *smem_slot;
}
template <int COLS>
__device__ __forceinline__ void tmem_free_cta(uint32_t base) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
TREX_FN void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
TREX_FN void tmem_ld4(tmem_t t, int col, vf (&out)[4]) {  // (results land asynchronously: read them only after tmem_wait4(out))
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(out[0]), "=f"(out[1]), "=f"(out[2]), "=f"(out[3]) : "r"(t.base + (uint32_t)col));
}
TREX_FN void tmem_wait4(vf (&v)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3])::"memory");
}

// 4 consecutive floats per lane (16-byte aligned offset): one 128-bit access
TREX_FN void ld4(const float* p, vi idx, vf out[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p + idx);
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
TREX_FN void ld4_if(const float* p, vi idx, vb pred, vf out[4]) {
  float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (pred) t = *reinterpret_cast<const float4*>(p + idx);
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
TREX_FN void st4_if(float* p, vi idx, const vf v[4], vb pred) {
  if (pred) *reinterpret_cast<float4*>(p + idx) = make_float4(v[0], v[1], v[2], v[3]);
}

// 16 consecutive floats (64-byte aligned offset): per-lane store / uniform load as four 128-bit accesses
TREX_FN void st16(float* p, vi idx, const vf v[16]) {
  float4* q = reinterpret_cast<float4*>(p + idx);
  q[0] = make_float4(v[0], v[1], v[2], v[3]); q[1] = make_float4(v[4], v[5], v[6], v[7]);
  q[2] = make_float4(v[8], v[9], v[10], v[11]); q[3] = make_float4(v[12], v[13], v[14], v[15]);
}
TREX_FN void ldu16(const float* p, int idx, float out[16]) {
  const float4* q = reinterpret_cast<const float4*>(p + idx);
  TREX_UNROLL for (int i = 0; i < 4; i++) { const float4 t = q[i]; out[4 * i] = t.x; out[4 * i + 1] = t.y; out[4 * i + 2] = t.z; out[4 * i + 3] = t.w; }
}
