// emu_main.cpp -- builds libtrex_emu.so: the PRODUCT kernel source (trex_core.h) executed on a CPU
// through the host lane emulator.  TEST INFRASTRUCTURE ONLY (no GPU in the build container).
#include "lane_emu.h"
#include "../../trex_gym_b200/csrc/trex_core.h"
#include "../../trex_gym_b200/csrc/trex_model.h"

#include <new>
#include <thread>

pthread_barrier_t* g_emu_cta_barrier = nullptr;

struct Emu {
  trex_host::ModelTables T;
  trex::Uniform P;
  trex::WarpShared S;            // the slab of a lone warp
  trex::WarpShared slabs[4];     // the slabs of a 4-warp CTA (packed inward pass)
  alignas(16) float scratch2[TREX_SOLVE2_SCRATCH];  // shared scratch of a solve2 warp
  float tmem[32][512];           // tensor memory of a solve2<true> warp (32 lanes x 512 columns)
  int heavy_tmem = 0;            // solve2 with the Delassus matrices in (emulated) tensor memory
  int solve_tmem = 0;            // solve4<TREX_KC, true>: Delassus blocks and sweep responses in (emulated) tensor memory
  int packed = 0;                // emu_step4 with n == 4: run the front phase as a 4-warp CTA (host threads)
  alignas(16) float work[4 * TREX_WORK_STRIDE];
  alignas(16) float workh[4 * TREX_HEAVY_STRIDE];
  alignas(16) float scratch[TREX_SOLVE_SCRATCH(TREX_KC)];
  int deferred = 1;
  int pack_reverse = 0;  // tests: fill the solver's lane groups from the top
  long long solves[5] = {0, 0, 0, 0, 0};  // substeps finished in front_phase / by solve4<0> / by solve4<TREX_KC>; [3] substep rounds run as a 4-warp CTA; [4] substeps finished by solve2
};

static_assert(trex::F_COUNT == 32, "field table");

extern "C" {
static char g_err[256];
const char* emu_last_error() { return g_err; }

void* emu_create(const void* blob, size_t bytes, int n_sub, float wd, float we, float wk, int max_episode_steps, int contacts,
                 int reset_mode, unsigned seed) {
  Emu* e = new Emu();
  if (!trex_host::build_tables(blob, bytes, e->T, trex::F_COUNT, trex::IF_COUNT)) {
    snprintf(g_err, sizeof g_err, "%s", e->T.err.c_str());
    delete e;
    return nullptr;
  }
  if ((int)e->T.params[trex_host::P_MAX_CONTACTS] != TREX_KMAX) {
    snprintf(g_err, sizeof g_err, "model max_contacts != TREX_KMAX");
    delete e;
    return nullptr;
  }
  trex_host::EnvConfig C;
  C.num_substeps = n_sub; C.distance_weight = wd; C.energy_weight = we; C.drift_weight = wk;
  C.max_episode_steps = max_episode_steps; C.enable_contacts = contacts;
  C.reset_mode = reset_mode; C.seed = seed;
  trex_host::fill_uniform(e->T, C, e->P);
  // shared memory is NOT zero on the device: poison the emulated slabs (all-ones = NaN floats, -1 ints) so that any read
  // of a never-written location shows up in the parity tests
  memset(&e->S, 0xff, sizeof(e->S));
  memset(e->slabs, 0xff, sizeof(e->slabs));
  memset(e->scratch, 0xff, sizeof(e->scratch));
  memset(e->work, 0xff, sizeof(e->work));
  memset(e->workh, 0xff, sizeof(e->workh));
  memset(e->scratch2, 0xff, sizeof(e->scratch2));
  memset(e->tmem, 0xff, sizeof(e->tmem));
  return e;
}
void emu_destroy(void* h) { delete (Emu*)h; }
int emu_state_stride() { return TREX_STATE_STRIDE; }
int emu_shared_bytes() { return (int)sizeof(trex::WarpShared); }

// one env step (or reset when force_reset) on n (1..4) consecutive environment records: the same phase sequence
// the library launches as kernels (front per environment, solve4 per four environments, tail per environment)
void emu_step4(void* h, int n, float* rec, const float* action, float* obs, float* reward, uint8_t* done, float* aux, int force_reset,
               long long env_id0) {
  Emu* e = (Emu*)h;
  const float* mdl = e->T.mdl.data();
  const int* mdli = e->T.mdli.data();
  const float* tasks = e->T.tasks.data();
  const float* cp = e->T.cand_p.data();
  const int* cl = e->T.cand_lane.data();
  if (!force_reset) {
    for (int r = 0; r < e->P.n_sub; r++) {
      // deferred environments are packed into the solver's lane groups in list order (any order is equivalent)
      // (one list per class, as in the library: contact-free, 1, 2, 3-4, 5-8 and more contacts)
      int envs[TREX_NCLASS][4] = {}, cnt[TREX_NCLASS] = {};
      int dres[4] = {0, 0, 0, 0};
      if (e->packed && n == 4) {
        // the library's 4-warp CTA: warps are host threads, cta_sync() a pthread barrier, warp 0 runs inward_packed
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, 4);
        g_emu_cta_barrier = &bar;
        e->solves[3]++;
        std::thread th[4];
        for (int i = 0; i < 4; i++)
          th[i] = std::thread([=, &dres]() {
            dres[i] = 255 & trex::front_phase<true>(e->P, mdl, mdli, tasks, cp, cl, e->slabs[i], rec + i * TREX_STATE_STRIDE,
                                              e->deferred ? e->work + i * TREX_WORK_STRIDE : nullptr, action + i * trex::NJ, r == 0,
                                              e->slabs, i, 0xf, e->deferred ? e->workh + i * TREX_HEAVY_STRIDE : nullptr);
          });
        for (int i = 0; i < 4; i++) th[i].join();
        g_emu_cta_barrier = nullptr;
        pthread_barrier_destroy(&bar);
      } else {
        for (int i = 0; i < n; i++)
          dres[i] = 255 & trex::front_phase(e->P, mdl, mdli, tasks, cp, cl, e->S, rec + i * TREX_STATE_STRIDE,
                                      e->deferred ? e->work + i * TREX_WORK_STRIDE : nullptr, action + i * trex::NJ, r == 0, nullptr, 0, 0,
                                      e->deferred ? e->workh + i * TREX_HEAVY_STRIDE : nullptr);
      }
      for (int i = 0; i < n; i++) {
        const int d = dres[i];
        e->solves[d == 0 ? 0 : (d == 1 ? 1 : (d < 1 + TREX_CLASS_HEAVY ? 2 : 4))]++;
        if (d) { int& c = cnt[d - 1]; envs[d - 1][e->pack_reverse ? 3 - c : c] = i; c++; }
      }
      for (int d = 0; d < TREX_NCLASS; d++) {
        if (!cnt[d]) continue;
        const int pending = e->pack_reverse ? (((1 << cnt[d]) - 1) << (4 - cnt[d])) : ((1 << cnt[d]) - 1);
        if (d == 0) trex::solve_phase<0>(e->P, e->scratch, e->work, rec, envs[0], pending);
        else if (d < TREX_CLASS_HEAVY && e->solve_tmem) { tmem_t tm4 = {e->tmem}; trex::solve_phase<TREX_KC, true>(e->P, e->scratch, e->work, rec, envs[d], pending, tm4); }
        else if (d < TREX_CLASS_HEAVY) trex::solve_phase<TREX_KC>(e->P, e->scratch, e->work, rec, envs[d], pending);
        else {  // class 5: two environments per warp (solve2), packed from the list like the kernel does
          int hv[4], nh = 0;
          for (int g = 0; g < 4; g++)
            if (pending & (1 << g)) hv[nh++] = envs[d][g];
          for (int i = 0; i < nh; i += 2) {
            const bool two = i + 1 < nh;
            int pair[2] = {hv[i], two ? hv[i + 1] : 0};
            int pend = two ? 3 : 1;
            if (e->pack_reverse && !two) { pair[1] = pair[0]; pair[0] = 0; pend = 2; }  // tests: a lone environment in the upper lane group
            tmem_t tm = {e->tmem};
            if (e->heavy_tmem) trex::heavy_phase<true>(e->P, e->scratch2, tm, e->work, e->workh, rec, pair, pend);
            else trex::heavy_phase<false>(e->P, e->scratch2, tm, e->work, e->workh, rec, pair, pend);
          }
        }
      }
    }
  }
  for (int i = 0; i < n; i++)
    trex::tail_phase(e->P, mdl, mdli, tasks, cp, cl, e->S, rec + i * TREX_STATE_STRIDE, obs ? obs + i * 75 : nullptr,
                     reward ? reward + i : nullptr, done ? done + i : nullptr, aux ? aux + i * TREX_AUX_STRIDE : nullptr,
                     force_reset != 0, env_id0 + i);
}
void emu_step(void* h, float* rec, const float* action, float* obs, float* reward, uint8_t* done, float* aux, int force_reset,
              long long env_id) {
  emu_step4(h, 1, rec, action, obs, reward, done, aux, force_reset, env_id);
}
void emu_solve_counts(void* h, long long* out) { for (int i = 0; i < 5; i++) out[i] = ((Emu*)h)->solves[i]; }
void emu_set_deferred(void* h, int on) {  // bit 0 deferral on, bit 1 reverse packing, bit 2 contact-free substeps only, bit 3 front phase as a 4-warp CTA, bit 4 more than TREX_KC contacts stay in the front phase
  Emu* e = (Emu*)h;
  e->heavy_tmem = (on >> 5) & 1;  // bit 5: solve2 keeps the Delassus matrices in tensor memory
  e->solve_tmem = (on >> 6) & 1;  // bit 6: solve4 keeps its Delassus blocks and sweep responses in tensor memory
  e->deferred = on & 1; e->pack_reverse = (on >> 1) & 1; e->P.defer_contacts = ((on >> 2) & 1) ? 0 : (((on >> 4) & 1) ? 1 : 2); e->packed = (on >> 3) & 1;
}
}
