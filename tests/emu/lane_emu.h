// lane_emu.h -- host emulation of the lane-vector vocabulary (TEST INFRASTRUCTURE ONLY).
//
// Counterpart of trex_gym_b200/csrc/lane_cuda.h: vf/vi/vb are 32-wide structs and the
// collectives are plain loops, so tests can execute the product kernel source
// (trex_gym_b200/csrc/trex_core.h) lane-for-lane on a CPU.  Never linked into the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define TREX_FN static inline
#define TREX_UNROLL
#define TREX_ROLLED
#define TREX_EMU 1

struct vb {
  bool v[32];
};
struct vi {
  int v[32];
  vi() {}
  vi(int x) { for (int l = 0; l < 32; l++) v[l] = x; }
};
struct vf {
  float v[32];
  vf() {}
  vf(float x) { for (int l = 0; l < 32; l++) v[l] = x; }
};

#define EMU_BINOP_F(op)                                                                            \
  TREX_FN vf operator op(const vf& a, const vf& b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b.v[l]; return r; } \
  TREX_FN vf operator op(const vf& a, float b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b; return r; }          \
  TREX_FN vf operator op(float a, const vf& b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = a op b.v[l]; return r; }
EMU_BINOP_F(+)
EMU_BINOP_F(-)
EMU_BINOP_F(*)
EMU_BINOP_F(/)
TREX_FN vf operator-(const vf& a) { vf r; for (int l = 0; l < 32; l++) r.v[l] = -a.v[l]; return r; }
TREX_FN vf& operator+=(vf& a, const vf& b) { for (int l = 0; l < 32; l++) a.v[l] += b.v[l]; return a; }
TREX_FN vf& operator-=(vf& a, const vf& b) { for (int l = 0; l < 32; l++) a.v[l] -= b.v[l]; return a; }
TREX_FN vf& operator*=(vf& a, const vf& b) { for (int l = 0; l < 32; l++) a.v[l] *= b.v[l]; return a; }

#define EMU_CMP_F(op)                                                                              \
  TREX_FN vb operator op(const vf& a, const vf& b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b.v[l]; return r; } \
  TREX_FN vb operator op(const vf& a, float b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b; return r; }          \
  TREX_FN vb operator op(float a, const vf& b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a op b.v[l]; return r; }
EMU_CMP_F(<)
EMU_CMP_F(>)
EMU_CMP_F(<=)
EMU_CMP_F(>=)
EMU_CMP_F(==)
EMU_CMP_F(!=)

#define EMU_BINOP_I(op)                                                                            \
  TREX_FN vi operator op(const vi& a, const vi& b) { vi r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b.v[l]; return r; } \
  TREX_FN vi operator op(const vi& a, int b) { vi r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b; return r; }            \
  TREX_FN vi operator op(int a, const vi& b) { vi r; for (int l = 0; l < 32; l++) r.v[l] = a op b.v[l]; return r; }
EMU_BINOP_I(+)
EMU_BINOP_I(-)
EMU_BINOP_I(*)
EMU_BINOP_I(&)
EMU_BINOP_I(|)
EMU_BINOP_I(>>)
EMU_BINOP_I(<<)
#define EMU_CMP_I(op)                                                                              \
  TREX_FN vb operator op(const vi& a, const vi& b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b.v[l]; return r; } \
  TREX_FN vb operator op(const vi& a, int b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] op b; return r; }
EMU_CMP_I(<)
EMU_CMP_I(>)
EMU_CMP_I(<=)
EMU_CMP_I(>=)
EMU_CMP_I(==)
EMU_CMP_I(!=)

TREX_FN vb operator&&(const vb& a, const vb& b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] && b.v[l]; return r; }
TREX_FN vb operator||(const vb& a, const vb& b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] || b.v[l]; return r; }
TREX_FN vb operator!(const vb& a) { vb r; for (int l = 0; l < 32; l++) r.v[l] = !a.v[l]; return r; }
TREX_FN vb operator&&(const vb& a, bool b) { vb r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] && b; return r; }
TREX_FN vb operator&&(bool a, const vb& b) { return b && a; }

TREX_FN vi lane_id() { vi r; for (int l = 0; l < 32; l++) r.v[l] = l; return r; }
TREX_FN vf vbroadcast(float x) { return vf(x); }
TREX_FN vf sel(const vb& p, const vf& a, const vf& b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
TREX_FN vf sel(const vb& p, const vf& a, float b) { return sel(p, a, vf(b)); }
TREX_FN vf sel(const vb& p, float a, const vf& b) { return sel(p, vf(a), b); }
TREX_FN vf sel(const vb& p, float a, float b) { return sel(p, vf(a), vf(b)); }
TREX_FN vi seli(const vb& p, const vi& a, const vi& b) { vi r; for (int l = 0; l < 32; l++) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
TREX_FN vf shfl(const vf& x, int src) { return vf(x.v[src & 31]); }
TREX_FN vf shflv(const vf& x, const vi& src) { vf r; for (int l = 0; l < 32; l++) r.v[l] = x.v[src.v[l] & 31]; return r; }
TREX_FN vf shfl_xor(const vf& x, int m) { vf r; for (int l = 0; l < 32; l++) r.v[l] = x.v[l ^ m]; return r; }
TREX_FN float lane_value(const vf& x, int lane) { return x.v[lane & 31]; }
TREX_FN int lane_value_i(const vi& x, int lane) { return x.v[lane & 31]; }
TREX_FN uint32_t vballot(const vb& p) { uint32_t m = 0; for (int l = 0; l < 32; l++) m |= (uint32_t)(p.v[l] ? 1u : 0u) << l; return m; }
TREX_FN bool vany(const vb& p) { return vballot(p) != 0; }
TREX_FN void warp_sync() {}

TREX_FN vf warp_sum(vf x) {
  for (int m = 16; m > 0; m >>= 1) x = x + shfl_xor(x, m);
  return x;
}
TREX_FN vf vmax(const vf& a, const vf& b);
TREX_FN vf warp_max(vf x) {
  for (int m = 16; m > 0; m >>= 1) x = vmax(x, shfl_xor(x, m));
  return x;
}

TREX_FN vf ld(const float* p, const vi& idx) { vf r; for (int l = 0; l < 32; l++) r.v[l] = p[idx.v[l]]; return r; }
TREX_FN vf ldg_ro(const float* p, const vi& idx) { return ld(p, idx); }
TREX_FN vf ld_if(const float* p, const vi& idx, const vb& pred, float dflt) { vf r; for (int l = 0; l < 32; l++) r.v[l] = pred.v[l] ? p[idx.v[l]] : dflt; return r; }
TREX_FN void st(float* p, const vi& idx, const vf& v) { for (int l = 0; l < 32; l++) p[idx.v[l]] = v.v[l]; }
TREX_FN void st_if(float* p, const vi& idx, const vf& v, const vb& pred) { for (int l = 0; l < 32; l++) if (pred.v[l]) p[idx.v[l]] = v.v[l]; }
TREX_FN void st_u8_if(uint8_t* p, const vi& idx, const vi& v, const vb& pred) { for (int l = 0; l < 32; l++) if (pred.v[l]) p[idx.v[l]] = (uint8_t)v.v[l]; }
TREX_FN float ldu(const float* p, int idx) { return p[idx]; }
TREX_FN int ldui(const int* p, int idx) { return p[idx]; }

TREX_FN vf vfma(const vf& a, const vf& b, const vf& c) { vf r; for (int l = 0; l < 32; l++) r.v[l] = fmaf(a.v[l], b.v[l], c.v[l]); return r; }
TREX_FN vf vfma(float a, const vf& b, const vf& c) { return vfma(vf(a), b, c); }
TREX_FN vf vfma(const vf& a, float b, const vf& c) { return vfma(a, vf(b), c); }
TREX_FN vf vfma(const vf& a, const vf& b, float c) { return vfma(a, b, vf(c)); }
TREX_FN vf vsqrt(const vf& x) { vf r; for (int l = 0; l < 32; l++) r.v[l] = sqrtf(x.v[l]); return r; }
TREX_FN vf vabs(const vf& x) { vf r; for (int l = 0; l < 32; l++) r.v[l] = fabsf(x.v[l]); return r; }
TREX_FN vf vmin(const vf& a, const vf& b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = fminf(a.v[l], b.v[l]); return r; }
TREX_FN vf vmax(const vf& a, const vf& b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = fmaxf(a.v[l], b.v[l]); return r; }
TREX_FN vf vmin(const vf& a, float b) { return vmin(a, vf(b)); }
TREX_FN vf vclamp_sym(const vf& x, const vf& m) {  // min.xorsign.abs: min(|x|, |m|) with sign(x) ^ sign(m)
  vf r;
  for (int l = 0; l < 32; l++) r.v[l] = (signbit(x.v[l]) != signbit(m.v[l]) ? -1.0f : 1.0f) * fminf(fabsf(x.v[l]), fabsf(m.v[l]));
  return r;
}
// packed FP32 pairs (device: fma.rn.f32x2 / mul.rn.f32x2, the same bits as two scalar operations)
TREX_FN void vfma2s(vf& c0, vf& c1, const vf& a0, const vf& a1, const vf& s) {
  for (int l = 0; l < 32; l++) { c0.v[l] = fmaf(a0.v[l], s.v[l], c0.v[l]); c1.v[l] = fmaf(a1.v[l], s.v[l], c1.v[l]); }
}
TREX_FN void vfma2v(vf& c0, vf& c1, const vf& a0, const vf& a1, const vf& b0, const vf& b1) {
  for (int l = 0; l < 32; l++) { c0.v[l] = fmaf(a0.v[l], b0.v[l], c0.v[l]); c1.v[l] = fmaf(a1.v[l], b1.v[l], c1.v[l]); }
}
TREX_FN void vmul2s(vf& r0, vf& r1, const vf& a0, const vf& a1, const vf& s) {
  for (int l = 0; l < 32; l++) { r0.v[l] = a0.v[l] * s.v[l]; r1.v[l] = a1.v[l] * s.v[l]; }
}
TREX_FN vf vclamp_sym(const vf& x, float m) { return vclamp_sym(x, vf(m)); }
TREX_FN vf vmax(const vf& a, float b) { return vmax(a, vf(b)); }
TREX_FN vf vsin(const vf& x) { vf r; for (int l = 0; l < 32; l++) r.v[l] = sinf(x.v[l]); return r; }
TREX_FN vf vcos(const vf& x) { vf r; for (int l = 0; l < 32; l++) r.v[l] = cosf(x.v[l]); return r; }
TREX_FN vf vdiv(const vf& a, const vf& b) { vf r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] / b.v[l]; return r; }
TREX_FN vf vdiv(float a, const vf& b) { return vdiv(vf(a), b); }
TREX_FN vf vdiv(const vf& a, float b) { return vdiv(a, vf(b)); }
TREX_FN vb visnan(const vf& x) { vb r; for (int l = 0; l < 32; l++) r.v[l] = !(fabsf(x.v[l]) <= 3.0e38f); return r; }
TREX_FN vi vf2i_bits(const vf& x) { vi r; for (int l = 0; l < 32; l++) memcpy(&r.v[l], &x.v[l], 4); return r; }
TREX_FN vf vi2f_bits(const vi& x) { vf r; for (int l = 0; l < 32; l++) memcpy(&r.v[l], &x.v[l], 4); return r; }
TREX_FN vf vi2f(const vi& x) { vf r; for (int l = 0; l < 32; l++) r.v[l] = (float)x.v[l]; return r; }

TREX_FN vi ldi(const int* p, const vi& idx) { vi r; for (int l = 0; l < 32; l++) r.v[l] = p[idx.v[l]]; return r; }
TREX_FN void sti_if(int* p, const vi& idx, const vi& v, const vb& pred) { for (int l = 0; l < 32; l++) if (pred.v[l]) p[idx.v[l]] = v.v[l]; }
TREX_FN vi rank_below(uint32_t mask) { vi r; for (int l = 0; l < 32; l++) r.v[l] = __builtin_popcount(mask & ((1u << l) - 1u)); return r; }
TREX_FN int popc_u(uint32_t m) { return __builtin_popcount(m); }
// compile the emulator with -ffp-contract=off so these stay un-fused
TREX_FN float fmul_rn(float a, float b) { return a * b; }
TREX_FN float fadd_rn(float a, float b) { return a + b; }
TREX_FN vf vmul_rn(const vf& a, const vf& b) { return a * b; }
TREX_FN int ctz_u(uint32_t m) { return __builtin_ctz(m); }
TREX_FN int clz_u(uint32_t m) { return __builtin_clz(m); }
TREX_FN vf vrsqrt(const vf& x) { vf r; for (int l = 0; l < 32; l++) r.v[l] = 1.0f / sqrtf(x.v[l]); return r; }
TREX_FN void stb(unsigned char* p, const vi& idx, const vi& v) { for (int l = 0; l < 32; l++) p[idx.v[l]] = (unsigned char)v.v[l]; }
TREX_FN vi ldb(const unsigned char* p, const vi& idx) { vi r; for (int l = 0; l < 32; l++) r.v[l] = (int)p[idx.v[l]]; return r; }
TREX_FN long long cycle_count() { return 0; }
// CTA barrier: the emulated warps of a CTA are host threads (emu_main.cpp) meeting at a pthread barrier
#include <pthread.h>
extern pthread_barrier_t* g_emu_cta_barrier;
TREX_FN void cta_sync() { if (g_emu_cta_barrier) pthread_barrier_wait(g_emu_cta_barrier); }

TREX_FN void philox4_uniform(const vi& c0, const vi& c1, const vi& c2, const vi& c3, uint32_t k0in, uint32_t k1in, vf out[4]) {
  for (int l = 0; l < 32; l++) {
    uint32_t c[4] = {(uint32_t)c0.v[l], (uint32_t)c1.v[l], (uint32_t)c2.v[l], (uint32_t)c3.v[l]};
    uint32_t k0 = k0in, k1 = k1in;
    for (int r = 0; r < 10; r++) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
      c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    for (int k = 0; k < 4; k++) out[k].v[l] = (float)(c[k] >> 8) * (1.0f / 16777216.0f);
  }
}

TREX_FN vi vmini(const vi& a, int b) { vi r; for (int l = 0; l < 32; l++) r.v[l] = a.v[l] < b ? a.v[l] : b; return r; }
TREX_FN vi vf2i(const vf& x) { vi r; for (int l = 0; l < 32; l++) r.v[l] = (int)x.v[l]; return r; }
TREX_FN vi warp_maxi(const vi& x) { int m = x.v[0]; for (int l = 1; l < 32; l++) m = x.v[l] > m ? x.v[l] : m; return vi(m); }
TREX_FN vf shfl_group8(const vf& x, int src) { vf r; for (int l = 0; l < 32; l++) r.v[l] = x.v[(l & ~7) | (src & 7)]; return r; }
TREX_FN vf shflv_group8(const vf& x, const vi& src) { vf r; for (int l = 0; l < 32; l++) r.v[l] = x.v[(l & ~7) | (src.v[l] & 7)]; return r; }
TREX_FN vf group8_sum(vf x) { for (int m = 4; m > 0; m >>= 1) x = x + shfl_xor(x, m); return x; }
TREX_FN vf group8_max(vf x) { for (int m = 4; m > 0; m >>= 1) x = vmax(x, shfl_xor(x, m)); return x; }

// tensor memory as a per-lane scratchpad (device: tcgen05.st / tcgen05.ld 32x32b.x4): here 32 lanes x 512 columns of host memory
struct tmem_t { float (*m)[512]; };
TREX_FN void tmem_st4(tmem_t t, int col, const vf& a, const vf& b, const vf& c, const vf& d) {
  for (int l = 0; l < 32; l++) { t.m[l][col] = a.v[l]; t.m[l][col + 1] = b.v[l]; t.m[l][col + 2] = c.v[l]; t.m[l][col + 3] = d.v[l]; }
}
TREX_FN void tmem_st_wait() {}
TREX_FN void tmem_ld4(tmem_t t, int col, vf (&out)[4]) { for (int l = 0; l < 32; l++) for (int k = 0; k < 4; k++) out[k].v[l] = t.m[l][col + k]; }
TREX_FN void tmem_wait4(vf (&)[4]) {}
TREX_FN vi shfl_xor_i(const vi& x, int m) { vi r; for (int l = 0; l < 32; l++) r.v[l] = x.v[l ^ m]; return r; }
TREX_FN vi sig_mix_v(const vi& h, const vi& w) { vi r; for (int l = 0; l < 32; l++) r.v[l] = (int)(((uint32_t)h.v[l] ^ (uint32_t)w.v[l]) * 16777619u); return r; }
TREX_FN vf shfl_group16(const vf& x, int src) { vf r; for (int l = 0; l < 32; l++) r.v[l] = x.v[(l & ~15) | (src & 15)]; return r; }
TREX_FN vf shflv_group16(const vf& x, const vi& src) { vf r; for (int l = 0; l < 32; l++) r.v[l] = x.v[(l & ~15) | (src.v[l] & 15)]; return r; }
TREX_FN vf group16_sum(vf x) { for (int m = 8; m > 0; m >>= 1) x = x + shfl_xor(x, m); return x; }
TREX_FN void ld2(const float* p, const vi& idx, vf out[2]) { for (int l = 0; l < 32; l++) for (int k = 0; k < 2; k++) out[k].v[l] = p[idx.v[l] + k]; }
TREX_FN void st2_if(float* p, const vi& idx, const vf v[2], const vb& pred) {
  for (int l = 0; l < 32; l++) if (pred.v[l]) for (int k = 0; k < 2; k++) p[idx.v[l] + k] = v[k].v[l];
}
TREX_FN void ld4(const float* p, const vi& idx, vf out[4]) { for (int l = 0; l < 32; l++) for (int k = 0; k < 4; k++) out[k].v[l] = p[idx.v[l] + k]; }
TREX_FN void ld4_if(const float* p, const vi& idx, const vb& pred, vf out[4]) {
  for (int l = 0; l < 32; l++) for (int k = 0; k < 4; k++) out[k].v[l] = pred.v[l] ? p[idx.v[l] + k] : 0.0f;
}
TREX_FN void st4_if(float* p, const vi& idx, const vf v[4], const vb& pred) {
  for (int l = 0; l < 32; l++) if (pred.v[l]) for (int k = 0; k < 4; k++) p[idx.v[l] + k] = v[k].v[l];
}
TREX_FN void st16(float* p, const vi& idx, const vf v[16]) { for (int l = 0; l < 32; l++) for (int k = 0; k < 16; k++) p[idx.v[l] + k] = v[k].v[l]; }
TREX_FN void ldu16(const float* p, int idx, float out[16]) { for (int k = 0; k < 16; k++) out[k] = p[idx + k]; }
