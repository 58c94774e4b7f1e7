// emu_main.cpp -- builds libtrex_emu.so: the PRODUCT kernel source (trex_core.h) executed on a CPU
// through the host lane emulator.  TEST INFRASTRUCTURE ONLY (no GPU in the build container).
#include "lane_emu.h"
#include "../../trex_gym_b200/csrc/trex_core.h"
#include "../../trex_gym_b200/csrc/trex_model.h"

#include <new>

struct Emu {
  trex_host::ModelTables T;
  trex::Uniform P;
  trex::WarpShared S;
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
  trex_host::EnvConfig C;
  C.num_substeps = n_sub; C.distance_weight = wd; C.energy_weight = we; C.drift_weight = wk;
  C.max_episode_steps = max_episode_steps; C.enable_contacts = contacts;
  C.reset_mode = reset_mode; C.seed = seed;
  trex_host::fill_uniform(e->T, C, e->P);
  memset(&e->S, 0, sizeof(e->S));
  return e;
}
void emu_destroy(void* h) { delete (Emu*)h; }
int emu_state_stride() { return TREX_STATE_STRIDE; }
int emu_shared_bytes() { return (int)sizeof(trex::WarpShared); }

// one env step (or reset when force_reset) on a single environment record
void emu_step(void* h, float* rec, const float* action, float* obs, float* reward, uint8_t* done, float* aux, int force_reset,
              long long env_id) {
  Emu* e = (Emu*)h;
  trex::env_step(e->P, e->T.mdl.data(), e->T.mdli.data(), e->T.tasks.data(), e->T.cand_p.data(), e->T.cand_lane.data(), e->S, rec,
                 action, obs, reward, done, aux, force_reset != 0, env_id);
}
}
