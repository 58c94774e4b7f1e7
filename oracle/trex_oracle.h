/*
 * trex_oracle.h -- CPU oracle for the trex-gym hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in pybullet (Bullet3
 * btMultiBody), which is not vendored under /root/reference, is unpinned
 * (setup.py:12) and is not installable here; the reference holds no golden
 * vectors for this path.  This file restates the published
 * Bullet algorithm from knowledge of its sources and is pinned only by analytic
 * known-answer tests (tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (trex_gym_b200) never does.
 */
#ifndef TREX_ORACLE_H
#define TREX_ORACLE_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct trex_oracle trex_oracle;

/* state vector layout (doubles):
 *   [0:3]   base COM position (world)
 *   [3:7]   base orientation quaternion x,y,z,w (base -> world), as pybullet reports it
 *   [7:10]  base angular velocity (world)
 *   [10:13] base linear velocity (world)
 *   [13:38] joint positions, revolute joints in pybullet link order
 *   [38:63] joint velocities
 *   [63:88] applied motor torque of the last substep
 *   [88:88+n_cand] cached normal impulse per contact candidate (warm start)   */
#define TREX_ORACLE_CORE_DIM 63

trex_oracle* trex_oracle_create(const void* blob, size_t bytes);
void trex_oracle_destroy(trex_oracle* o);
const char* trex_oracle_last_error(void);

int trex_oracle_state_dim(const trex_oracle* o);
int trex_oracle_num_candidates(const trex_oracle* o);
void trex_oracle_get_state(const trex_oracle* o, double* out);
void trex_oracle_set_state(trex_oracle* o, const double* in);

/* num_substeps n: dt = time_step/n, iterations = (int)(solver_iterations/n)  (trex_env.py:71-73) */
void trex_oracle_set_substeps(trex_oracle* o, int n);
void trex_oracle_set_reward_weights(trex_oracle* o, double distance, double energy, double drift);
void trex_oracle_enable_contacts(trex_oracle* o, int on);
/* btMultiBody's fixed-base mode (m_fixedBase): the base neither accelerates nor responds to impulses.  The reference loads
 * the T-rex with a floating base (trex_robot.py:47-56); this switch exists for the fixed-base pendulum known-answer test. */
void trex_oracle_set_fixed_base(trex_oracle* o, int on);

/* TrexBulletEnv.reset (trex_env.py:98-122): reset pose, zero-force motors, ONE physics step. obs75 may be NULL */
void trex_oracle_reset(trex_oracle* o, double* obs75);
/* TrexBulletEnv.step (trex_env.py:128-154). action in name-sorted joint order. */
void trex_oracle_step(trex_oracle* o, const double* action25, double* obs75, double* reward);
/* run T env steps; actions [T][25]; obs_out [T][75] and reward_out [T] may be NULL */
void trex_oracle_run(trex_oracle* o, const double* actions, int T, double* obs_out, double* reward_out);

/* one pybullet stepSimulation with explicit motor settings (dof order): target, max impulse per motor */
void trex_oracle_substep(trex_oracle* o, const double* target25, double max_impulse);

/* diagnostics */
void trex_oracle_head_position(trex_oracle* o, double* xyz);
void trex_oracle_reward_terms(const trex_oracle* o, double* three);
int trex_oracle_last_iterations(const trex_oracle* o);   /* PGS iterations of the last substep */
int trex_oracle_last_num_contacts(const trex_oracle* o); /* contact points with rows in the last substep */
int trex_oracle_last_num_limit_rows(const trex_oracle* o);
long trex_oracle_total_iterations(const trex_oracle* o);
long trex_oracle_total_substeps(const trex_oracle* o);
/* Active-set signature of the last trex_oracle_step (24-bit hash over, per substep: the contact candidates with rows, the
 * limit / normal rows that ended with a positive impulse, the motors / friction pairs that ended on their bound, the PGS
 * iterations executed) -- defined exactly as record slot 158 of the kernels (trex_gym_b200/csrc/trex_core.h: StepStats), so
 * equal signatures mean kernel and oracle solved the same complementarity problem.  words: the last substep's six words. */
unsigned trex_oracle_signature(const trex_oracle* o);
void trex_oracle_reset_signature(trex_oracle* o);
void trex_oracle_signature_words(const trex_oracle* o, unsigned* out6);
/* total linear momentum (3), angular momentum about world origin (3), kinetic energy (1), total mass (1), COM (3) */
void trex_oracle_momentum(trex_oracle* o, double* out11);
/* joint-space inverse mass matrix column via the unit-impulse pass (for tests): out[31] */
void trex_oracle_minv_column(trex_oracle* o, int dof, double* out31);
/* PGS residual (max (dlambda/jacDiagABInv)^2) after each iteration of the last substep */
void trex_oracle_residual_history(const trex_oracle* o, double* out, int n);
/* world position of contact candidate k */
void trex_oracle_candidate_position(trex_oracle* o, int k, double* xyz);

#ifdef __cplusplus
}
#endif
#endif
