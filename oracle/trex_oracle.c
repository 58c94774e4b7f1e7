/*
 * trex_oracle.c -- double-precision CPU restatement of the reference hot path:
 *   TrexBulletEnv.step / reset          (/root/reference/trex_gym/trex_env.py:98-154)
 *   TrexRobot.set_actions / get_observations / get_head_position / get_total_joint_power
 *                                        (/root/reference/trex_gym/trex_robot.py:300-422)
 *   pybullet.stepSimulation              (external, called at trex_env.py:120,150)
 *
 * TEST INFRASTRUCTURE ONLY -- see trex_oracle.h.  PARITY UNPINNED: pybullet (Bullet3,
 * version unpinned, setup.py:12) is absent from /root/reference and from this image; the
 * stepSimulation part below restates Bullet's published btMultiBody pipeline
 * (btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof,
 *  btMultiBody::calcAccelerationDeltasMultiDof, btMultiBodyJointMotor,
 *  btMultiBodyJointLimitConstraint, btMultiBodyConstraintSolver, btMultiBody::stepPositionsMultiDof,
 *  PhysicsServerCommandProcessor joint damping) as summarised in SURVEY.md Appendix A.
 *
 * It simulates the FULL pybullet multibody: floating base + 132 links in pybullet link
 * order, fixed joints kept as 0-DoF links, every link frame = its URDF inertial frame --
 * deliberately a different formulation from the CUDA kernels (26 merged bodies, joint
 * frames, float), so that agreement between the two is evidence.
 */
#include "trex_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 160
#define MAXDOF 32
#define MAXU (MAXDOF + 6)
#define MAXC 64
#define MAXROWS (2 * MAXDOF + 3 * MAXC)

static __thread char g_err[256];
const char* trex_oracle_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ small algebra */
typedef double v3[3];
typedef double m3[9]; /* row major */
typedef double sv[6]; /* spatial: [0:3] angular/torque, [3:6] linear/force */

static inline void v3set(v3 a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static inline void v3cpy(v3 a, const v3 b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static inline double v3dot(const v3 a, const v3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3cross(v3 o, const v3 a, const v3 b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline double v3norm(const v3 a) { return sqrt(v3dot(a, a)); }
static inline void m3mulv(v3 o, const m3 M, const v3 a) {
  double x = M[0] * a[0] + M[1] * a[1] + M[2] * a[2];
  double y = M[3] * a[0] + M[4] * a[1] + M[5] * a[2];
  double z = M[6] * a[0] + M[7] * a[1] + M[8] * a[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3tmulv(v3 o, const m3 M, const v3 a) { /* M^T a */
  double x = M[0] * a[0] + M[3] * a[1] + M[6] * a[2];
  double y = M[1] * a[0] + M[4] * a[1] + M[7] * a[2];
  double z = M[2] * a[0] + M[5] * a[1] + M[8] * a[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3mul(m3 O, const m3 A, const m3 B) {
  m3 T;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
  memcpy(O, T, sizeof(m3));
}
/* rotation matrix of angle about unit axis (active rotation) */
static void m3axisangle(m3 R, const v3 a, double th) {
  double c = cos(th), s = sin(th), t = 1 - c;
  R[0] = t * a[0] * a[0] + c;        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c;        R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
}
/* quaternion (x,y,z,w) -> rotation matrix (rotates body coords into world coords) */
static void quat2mat(m3 R, const double q[4]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double n = x * x + y * y + z * z + w * w, s = 2.0 / n;
  R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - w * z);     R[2] = s * (x * z + w * y);
  R[3] = s * (x * y + w * z);     R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - w * x);
  R[6] = s * (x * z - w * y);     R[7] = s * (y * z + w * x);     R[8] = 1 - s * (x * x + y * y);
}
static inline double svdot(const sv a, const sv b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

/* spatial transform parent -> child: rotation R (parent coords -> child coords) and
 * r = vector from parent origin to child origin, expressed in CHILD coordinates
 * (btSpatialTransformationMatrix: m_rotMat, m_trnVec). */
static void xf_motion(sv out, const m3 R, const v3 r, const sv in) { /* transform() */
  v3 a, l, t;
  m3mulv(a, R, in);
  m3mulv(l, R, in + 3);
  v3cross(t, r, a);
  out[0] = a[0]; out[1] = a[1]; out[2] = a[2];
  out[3] = l[0] - t[0]; out[4] = l[1] - t[1]; out[5] = l[2] - t[2];
}
static void xf_force_inv(sv out, const m3 R, const v3 r, const sv in) { /* transformInverse(): child -> parent */
  v3 t, n, f;
  v3cross(t, r, in + 3);
  t[0] += in[0]; t[1] += in[1]; t[2] += in[2];
  m3tmulv(n, R, t);
  m3tmulv(f, R, in + 3);
  out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
  out[3] = f[0]; out[4] = f[1]; out[5] = f[2];
}
/* 6x6 motion transform matrix X (v_child = X v_parent) */
static void xf_matrix(double X[36], const m3 R, const v3 r) {
  memset(X, 0, 36 * sizeof(double));
  double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
  m3 rxR;
  m3mul(rxR, rx, R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      X[6 * i + j] = R[3 * i + j];
      X[6 * (i + 3) + (j + 3)] = R[3 * i + j];
      X[6 * (i + 3) + j] = -rxR[3 * i + j];
    }
}
/* P += X^T I X */
static void inertia_to_parent_add(double P[36], const double X[36], const double I[36]) {
  double T[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += I[6 * i + k] * X[6 * k + j];
      T[6 * i + j] = s;
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += X[6 * k + i] * T[6 * k + j];
      P[6 * i + j] += s;
    }
}
static void mat6mulv(sv o, const double M[36], const sv a) {
  sv t;
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += M[6 * i + k] * a[k];
    t[i] = s;
  }
  memcpy(o, t, sizeof(sv));
}
/* solve A x = b for symmetric positive definite 6x6 (Gaussian elimination with partial pivoting) */
static void solve6(const double A[36], const sv b, sv x) {
  double M[6][7];
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) M[i][j] = A[6 * i + j];
    M[i][6] = b[i];
  }
  for (int c = 0; c < 6; c++) {
    int p = c;
    for (int r = c + 1; r < 6; r++)
      if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 7; j++) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
    double inv = 1.0 / M[c][c];
    for (int r = c + 1; r < 6; r++) {
      double f = M[r][c] * inv;
      for (int j = c; j < 7; j++) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 5; i >= 0; i--) {
    double s = M[i][6];
    for (int j = i + 1; j < 6; j++) s -= M[i][j] * x[j];
    x[i] = s / M[i][i];
  }
}

/* ------------------------------------------------------------------ model + state */
enum {
  P_TIME_STEP, P_SOLVER_ITERS, P_NUM_SUBSTEPS, P_GRAVITY, P_KP, P_KD, P_MAX_TORQUE, P_LIN_DAMP, P_ANG_DAMP,
  P_MAX_COORD_VEL, P_ERP, P_CONTACT_ERP, P_SPLIT_THRESH, P_LINEAR_SLOP, P_RESIDUAL, P_WARMSTART, P_FRICTION,
  P_BREAKING, P_FLOOR, P_LIMIT_MAX_IMPULSE, P_RESET_HEIGHT, P_TARGET_HEIGHT, P_MAX_CONTACTS, P_COUNT
};

typedef struct {
  double J[MAXU], dV[MAXU];
  double jdi, rhs, lo, hi, lam, mu;
  int fric_index; /* friction rows: index of the normal row */
  int cand;       /* contact rows: candidate index */
  int motor_dof;  /* motor rows: dof, else -1 */
  int limit_dof;  /* joint-limit rows: dof, else -1 */
} row_t;

struct trex_oracle {
  int n, ndof, nu;
  int parent[MAXL], jtype[MAXL], dof[MAXL], dof_link[MAXDOF];
  double mass[MAXL + 1], inertia[MAXL + 1][3];
  m3 rot0[MAXL];
  v3 axis[MAXL], dvec[MAXL], evec[MAXL];
  double lower[MAXL], upper[MAXL], damping[MAXL], start_q[MAXL];
  int obs_dof[MAXDOF], head_link;
  int ncand, cand_link[MAXC];
  v3 cand_local[MAXC];
  double cand_r[MAXC]; /* sphere candidates (contact primitives): radius, 0 = a point */
  int n_order, nc_order[2 * MAXDOF];
  double P[P_COUNT];
  int n_sub, contacts_on;
  int fixed_base; /* btMultiBody m_fixedBase: base acceleration and acceleration deltas forced to zero (KATs only) */
  double w_dist, w_energy, w_drift;
  /* state */
  v3 pos, omega, vel;
  double quat[4];
  double q[MAXDOF], qd[MAXDOF], tau[MAXDOF], lam_cache[MAXC];
  /* kinematics cache */
  m3 Rp[MAXL + 1], Rw[MAXL + 1]; /* rot_from_parent, rot_from_world (world coords -> link coords) */
  v3 rvec[MAXL], comw[MAXL + 1];
  /* ABA cache */
  sv S[MAXL], h[MAXL];
  double invD[MAXL];
  double IA0[36];
  /* diagnostics */
  int last_iters, last_contacts, last_limits;
  long total_iters, total_substeps;
  double rterm[3];
  double resid_hist[512];
  uint32_t sig;          /* running active-set signature (reset at the start of trex_oracle_step / on request) */
  uint32_t sig_words[6]; /* the last substep's words: K lo, K hi, L, M, N | F << 16, iterations */
  row_t* rows;
};

/* ---- blob parsing (format: trex_gym_b200/model_blob.py) ---- */
typedef struct { const uint8_t* base; size_t bytes; uint32_t nsec; } blob_t;
static int blob_find(const blob_t* b, const char* name, uint32_t want_dtype, const void** data, uint32_t* count) {
  for (uint32_t i = 0; i < b->nsec; i++) {
    const uint8_t* e = b->base + 16 + 40 * (size_t)i;
    char nm[25];
    memcpy(nm, e, 24);
    nm[24] = 0;
    if (strcmp(nm, name) == 0) {
      uint32_t dt, cnt;
      uint64_t off;
      memcpy(&dt, e + 24, 4); memcpy(&cnt, e + 28, 4); memcpy(&off, e + 32, 8);
      if (dt != want_dtype) { snprintf(g_err, sizeof g_err, "section %s: wrong dtype", name); return -1; }
      if (off + (size_t)cnt * (dt == 0 ? 8 : 4) > b->bytes) { snprintf(g_err, sizeof g_err, "section %s: out of range", name); return -1; }
      *data = b->base + off;
      *count = cnt;
      return 0;
    }
  }
  snprintf(g_err, sizeof g_err, "section %s missing", name);
  return -1;
}
#define GETF(name, ptr, cnt) do { const void* _d; if (blob_find(&b, name, 0, &_d, &cnt)) goto fail; ptr = (const double*)_d; } while (0)
#define GETI(name, ptr, cnt) do { const void* _d; if (blob_find(&b, name, 1, &_d, &cnt)) goto fail; ptr = (const int32_t*)_d; } while (0)

trex_oracle* trex_oracle_create(const void* blob, size_t bytes) {
  trex_oracle* o = (trex_oracle*)calloc(1, sizeof(trex_oracle));
  if (!o) return NULL;
  o->rows = (row_t*)calloc(MAXROWS, sizeof(row_t));
  blob_t b = {(const uint8_t*)blob, bytes, 0};
  if (bytes < 16 || memcmp(blob, "TREXMDL1", 8) != 0) { snprintf(g_err, sizeof g_err, "bad blob magic"); goto fail; }
  memcpy(&b.nsec, b.base + 12, 4);
  const double* f; const int32_t* ip; uint32_t c;
  GETI("full_n_links", ip, c); o->n = ip[0];
  if (o->n + 1 > MAXL) { snprintf(g_err, sizeof g_err, "too many links"); goto fail; }
  GETF("param_values", f, c);
  if (c != P_COUNT) { snprintf(g_err, sizeof g_err, "param count %u != %d", c, P_COUNT); goto fail; }
  memcpy(o->P, f, sizeof(double) * P_COUNT);
  GETI("full_parent", ip, c); for (int i = 0; i < o->n; i++) o->parent[i] = ip[i];
  GETI("full_jtype", ip, c); for (int i = 0; i < o->n; i++) o->jtype[i] = ip[i];
  GETI("full_dof", ip, c);
  o->ndof = 0;
  for (int i = 0; i < o->n; i++) { o->dof[i] = ip[i]; if (ip[i] >= 0) { o->dof_link[ip[i]] = i; o->ndof++; } }
  if (o->ndof > MAXDOF) { snprintf(g_err, sizeof g_err, "too many dofs"); goto fail; }
  o->nu = 6 + o->ndof;
  GETF("full_mass", f, c); memcpy(o->mass, f, sizeof(double) * (o->n + 1));
  GETF("full_inertia", f, c); memcpy(o->inertia, f, sizeof(double) * 3 * (o->n + 1));
  GETF("full_rot0", f, c); memcpy(o->rot0, f, sizeof(double) * 9 * o->n);
  GETF("full_axis", f, c); memcpy(o->axis, f, sizeof(double) * 3 * o->n);
  GETF("full_d", f, c); memcpy(o->dvec, f, sizeof(double) * 3 * o->n);
  GETF("full_e", f, c); memcpy(o->evec, f, sizeof(double) * 3 * o->n);
  GETF("full_lower", f, c); memcpy(o->lower, f, sizeof(double) * o->n);
  GETF("full_upper", f, c); memcpy(o->upper, f, sizeof(double) * o->n);
  GETF("full_damping", f, c); memcpy(o->damping, f, sizeof(double) * o->n);
  GETF("full_start_q", f, c); memcpy(o->start_q, f, sizeof(double) * o->n);
  GETI("full_head_link", ip, c); o->head_link = ip[0];
  GETI("obs_dof", ip, c); for (uint32_t i = 0; i < c; i++) o->obs_dof[i] = ip[i];
  GETI("full_cand_link", ip, c); o->ncand = (int)c;
  if (o->ncand > MAXC) { snprintf(g_err, sizeof g_err, "too many contact candidates"); goto fail; }
  for (int i = 0; i < o->ncand; i++) o->cand_link[i] = ip[i];
  GETF("full_cand_local", f, c); memcpy(o->cand_local, f, sizeof(double) * 3 * o->ncand);
  { /* optional section (older blobs have point candidates only) */
    const void* d_; uint32_t c_;
    if (blob_find(&b, "full_cand_r", 0, &d_, &c_) == 0 && (int)c_ == o->ncand) memcpy(o->cand_r, d_, sizeof(double) * o->ncand);
    g_err[0] = 0;
  }
  GETI("noncontact_order", ip, c); o->n_order = (int)c;
  for (uint32_t i = 0; i < c; i++) o->nc_order[i] = ip[i];
  o->n_sub = (int)o->P[P_NUM_SUBSTEPS];
  o->contacts_on = 1;
  o->w_dist = 1.0; o->w_energy = 0.005; o->w_drift = 0.002; /* trex_env.py:42-44 */
  o->quat[3] = 1.0;
  return o;
fail:
  free(o->rows);
  free(o);
  return NULL;
}
void trex_oracle_destroy(trex_oracle* o) { if (o) { free(o->rows); free(o); } }

int trex_oracle_state_dim(const trex_oracle* o) { return 63 + 25 + o->ncand; }
int trex_oracle_num_candidates(const trex_oracle* o) { return o->ncand; }
void trex_oracle_get_state(const trex_oracle* o, double* s) {
  memcpy(s, o->pos, 24); memcpy(s + 3, o->quat, 32); memcpy(s + 7, o->omega, 24); memcpy(s + 10, o->vel, 24);
  memcpy(s + 13, o->q, 8 * 25); memcpy(s + 38, o->qd, 8 * 25); memcpy(s + 63, o->tau, 8 * 25);
  memcpy(s + 88, o->lam_cache, 8 * o->ncand);
}
void trex_oracle_set_state(trex_oracle* o, const double* s) {
  memcpy(o->pos, s, 24); memcpy(o->quat, s + 3, 32); memcpy(o->omega, s + 7, 24); memcpy(o->vel, s + 10, 24);
  memcpy(o->q, s + 13, 8 * 25); memcpy(o->qd, s + 38, 8 * 25); memcpy(o->tau, s + 63, 8 * 25);
  memcpy(o->lam_cache, s + 88, 8 * o->ncand);
}
void trex_oracle_set_substeps(trex_oracle* o, int n) { o->n_sub = n < 1 ? 1 : n; }
void trex_oracle_set_reward_weights(trex_oracle* o, double d, double e, double k) { o->w_dist = d; o->w_energy = e; o->w_drift = k; }
void trex_oracle_enable_contacts(trex_oracle* o, int on) { o->contacts_on = on; }
void trex_oracle_set_fixed_base(trex_oracle* o, int on) { o->fixed_base = on; }
unsigned trex_oracle_signature(const trex_oracle* o) { return o->sig; }
void trex_oracle_reset_signature(trex_oracle* o) { o->sig = 0; }
void trex_oracle_signature_words(const trex_oracle* o, unsigned* out6) { for (int i = 0; i < 6; i++) out6[i] = o->sig_words[i]; }
int trex_oracle_last_iterations(const trex_oracle* o) { return o->last_iters; }
int trex_oracle_last_num_contacts(const trex_oracle* o) { return o->last_contacts; }
int trex_oracle_last_num_limit_rows(const trex_oracle* o) { return o->last_limits; }
long trex_oracle_total_iterations(const trex_oracle* o) { return o->total_iters; }
long trex_oracle_total_substeps(const trex_oracle* o) { return o->total_substeps; }
void trex_oracle_reward_terms(const trex_oracle* o, double* t) { t[0] = o->rterm[0]; t[1] = o->rterm[1]; t[2] = o->rterm[2]; }

/* ------------------------------------------------------------------ kinematics
 * btMultibodyLink::updateCacheMultiDof + btMultiBody::forwardKinematics:
 *   cachedRotParentToThis = quat(axis, -q) * zeroRotParentToThis
 *   cachedRVector         = rotate(cachedRot, eVector) + dVector                */
static void forward_kinematics(trex_oracle* o) {
  m3 Rbw;
  quat2mat(Rbw, o->quat); /* base -> world */
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o->Rw[0][3 * r + c] = Rbw[3 * c + r]; /* world -> base */
  memcpy(o->Rp[0], o->Rw[0], sizeof(m3));
  v3cpy(o->comw[0], o->pos);
  for (int i = 0; i < o->n; i++) {
    int p = o->parent[i] + 1;
    if (o->jtype[i] == 1) {
      m3 Rq;
      m3axisangle(Rq, o->axis[i], -o->q[o->dof[i]]);
      m3mul(o->Rp[i + 1], Rq, o->rot0[i]);
    } else {
      memcpy(o->Rp[i + 1], o->rot0[i], sizeof(m3));
    }
    v3 t;
    m3mulv(t, o->Rp[i + 1], o->evec[i]);
    o->rvec[i][0] = t[0] + o->dvec[i][0]; o->rvec[i][1] = t[1] + o->dvec[i][1]; o->rvec[i][2] = t[2] + o->dvec[i][2];
    m3mul(o->Rw[i + 1], o->Rp[i + 1], o->Rw[p]);
    m3tmulv(t, o->Rw[i + 1], o->rvec[i]);
    o->comw[i + 1][0] = o->comw[p][0] + t[0]; o->comw[i + 1][1] = o->comw[p][1] + t[1]; o->comw[i + 1][2] = o->comw[p][2] + t[2];
    if (o->jtype[i] == 1) { /* joint motion subspace: (axis, axis x d) */
      v3 ab;
      v3cross(ab, o->axis[i], o->dvec[i]);
      o->S[i][0] = o->axis[i][0]; o->S[i][1] = o->axis[i][1]; o->S[i][2] = o->axis[i][2];
      o->S[i][3] = ab[0]; o->S[i][4] = ab[1]; o->S[i][5] = ab[2];
    } else {
      memset(o->S[i], 0, sizeof(sv));
    }
  }
}

static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* btMultiBody::applyDeltaVeeMultiDof: u += m * dv, each coordinate clamped to +-maxCoordinateVelocity */
static void apply_delta_vee(trex_oracle* o, const double* dv, double mult) {
  double mx = o->P[P_MAX_COORD_VEL];
  for (int k = 0; k < 3; k++) o->omega[k] = clampd(o->omega[k] + dv[k] * mult, -mx, mx);
  for (int k = 0; k < 3; k++) o->vel[k] = clampd(o->vel[k] + dv[3 + k] * mult, -mx, mx);
  for (int k = 0; k < o->ndof; k++) o->qd[k] = clampd(o->qd[k] + dv[6 + k] * mult, -mx, mx);
}

/* ------------------------------------------------------------------ ABA
 * btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof.
 * joint_tau[dof] = explicit joint torques (pybullet joint damping, A.3 step 0).
 * Writes accelerations to acc[nu] = (omega_dot world, v_dot world, qdd) and caches S, h, invD, IA0. */
static void aba(trex_oracle* o, const double* joint_tau, double* acc) {
  const int n = o->n;
  static __thread sv vel[MAXL + 1], cor[MAXL], pA[MAXL + 1], a[MAXL + 1];
  static __thread double IA[MAXL + 1][36], Y[MAXL];
  const double k_lin = o->P[P_LIN_DAMP], k_ang = o->P[P_ANG_DAMP], g = o->P[P_GRAVITY];
  const v3 gw = {0, 0, -g};

  /* base */
  m3mulv(vel[0], o->Rw[0], o->omega);
  m3mulv(vel[0] + 3, o->Rw[0], o->vel);
  for (int b = 0; b <= n; b++) {
    if (b > 0) {
      int i = b - 1, p = o->parent[i] + 1;
      xf_motion(vel[b], o->Rp[b], o->rvec[i], vel[p]);
      sv sj = {0, 0, 0, 0, 0, 0};
      if (o->jtype[i] == 1) {
        double qd = o->qd[o->dof[i]];
        for (int k = 0; k < 6; k++) sj[k] = o->S[i][k] * qd;
        for (int k = 0; k < 6; k++) vel[b][k] += sj[k];
      }
      /* c = v x (S qd)  (btSpatialMotionVector::cross) */
      v3 t1, t2, t3;
      v3cross(t1, vel[b], sj);
      v3cross(t2, vel[b] + 3, sj);
      v3cross(t3, vel[b], sj + 3);
      cor[i][0] = t1[0]; cor[i][1] = t1[1]; cor[i][2] = t1[2];
      cor[i][3] = t2[0] + t3[0]; cor[i][4] = t2[1] + t3[1]; cor[i][5] = t2[2] + t3[2];
    }
    const double m = o->mass[b];
    const double* I = o->inertia[b];
    double* w = vel[b];
    double* vl = vel[b] + 3;
    /* external force: gravity m*g on every link, as a force (not a base acceleration) */
    v3 fw = {gw[0] * m, gw[1] * m, gw[2] * m}, fl;
    m3mulv(fl, o->Rw[b], fw);
    pA[b][0] = 0; pA[b][1] = 0; pA[b][2] = 0;
    pA[b][3] = -fl[0]; pA[b][4] = -fl[1]; pA[b][5] = -fl[2];
    /* damping: I w (k + k|w|), m v (k + k|v|) */
    double wn = v3norm(w), vn = v3norm(vl);
    for (int k = 0; k < 3; k++) {
      pA[b][k] += I[k] * w[k] * (k_ang + k_ang * wn);
      pA[b][3 + k] += m * vl[k] * (k_lin + k_lin * vn);
    }
    /* gyroscopic: w x (I w) ; m (w x v) */
    v3 Iw = {I[0] * w[0], I[1] * w[1], I[2] * w[2]}, gy, wv;
    v3cross(gy, w, Iw);
    v3cross(wv, w, vl);
    for (int k = 0; k < 3; k++) { pA[b][k] += gy[k]; pA[b][3 + k] += m * wv[k]; }
    memset(IA[b], 0, sizeof(IA[b]));
    IA[b][0] = I[0]; IA[b][7] = I[1]; IA[b][14] = I[2];
    IA[b][21] = m; IA[b][28] = m; IA[b][35] = m;
  }

  /* inward */
  for (int i = n - 1; i >= 0; i--) {
    int b = i + 1, p = o->parent[i] + 1;
    double Ia[36];
    sv pa;
    memcpy(Ia, IA[b], sizeof(Ia));
    if (o->jtype[i] == 1) {
      mat6mulv(o->h[i], IA[b], o->S[i]);
      double D = svdot(o->S[i], o->h[i]);
      Y[i] = joint_tau[o->dof[i]] - svdot(o->S[i], pA[b]) - svdot(cor[i], o->h[i]);
      o->invD[i] = (D >= 2.220446049250313e-16) ? 1.0 / D : 0.0; /* SIMD_EPSILON (double build) */
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) Ia[6 * r + c] -= o->h[i][r] * o->h[i][c] * o->invD[i];
      /* Zp += pXi * (Zi + Ii*ci + hi*Yi/Di): the FULL articulated inertia multiplies the Coriolis
       * term here because Y already carries -h.c (Mirtich's form, as in Bullet) */
      sv t;
      mat6mulv(t, IA[b], cor[i]);
      for (int k = 0; k < 6; k++) pa[k] = pA[b][k] + t[k] + o->h[i][k] * (o->invD[i] * Y[i]);
    } else {
      memset(o->h[i], 0, sizeof(sv));
      o->invD[i] = 0;
      Y[i] = 0;
      memcpy(pa, pA[b], sizeof(sv));
    }
    double X[36];
    xf_matrix(X, o->Rp[b], o->rvec[i]);
    inertia_to_parent_add(IA[p], X, Ia);
    sv pp;
    xf_force_inv(pp, o->Rp[b], o->rvec[i], pa);
    for (int k = 0; k < 6; k++) pA[p][k] += pp[k];
  }
  memcpy(o->IA0, IA[0], sizeof(o->IA0));

  /* base acceleration: a0 = -IA0^-1 pA0 */
  sv res;
  solve6(IA[0], pA[0], res);
  for (int k = 0; k < 6; k++) a[0][k] = o->fixed_base ? 0.0 : -res[k];

  /* outward */
  for (int i = 0; i < n; i++) {
    int b = i + 1, p = o->parent[i] + 1;
    xf_motion(a[b], o->Rp[b], o->rvec[i], a[p]);
    if (o->jtype[i] == 1) {
      double qdd = o->invD[i] * (Y[i] - svdot(a[b], o->h[i]));
      acc[6 + o->dof[i]] = qdd;
      for (int k = 0; k < 6; k++) a[b][k] += cor[i][k] + o->S[i][k] * qdd;
    }
  }
  /* back to world: omega_dot = R0^T a_ang ; v_dot = R0^T (a_lin + w x v) */
  v3 wv, t;
  v3cross(wv, vel[0], vel[0] + 3);
  m3tmulv(acc, o->Rw[0], a[0]);
  v3set(t, a[0][3] + wv[0], a[0][4] + wv[1], a[0][5] + wv[2]);
  m3tmulv(acc + 3, o->Rw[0], t);
  if (o->fixed_base) memset(acc, 0, 6 * sizeof(double));
}

/* btMultiBody::calcAccelerationDeltasMultiDof: out = M^-1 force, generalized coordinates
 * (base torque world, base force world, joint torques), using the h / invD / IA0 cached by aba(). */
static void accel_deltas(const trex_oracle* o, const double* force, double* out) {
  const int n = o->n;
  static __thread sv z[MAXL + 1], a[MAXL + 1];
  static __thread double Y[MAXL];
  memset(z, 0, sizeof(sv) * (n + 1));
  v3 t;
  m3mulv(t, o->Rw[0], force);
  z[0][0] = -t[0]; z[0][1] = -t[1]; z[0][2] = -t[2];
  m3mulv(t, o->Rw[0], force + 3);
  z[0][3] = -t[0]; z[0][4] = -t[1]; z[0][5] = -t[2];
  for (int i = n - 1; i >= 0; i--) {
    int b = i + 1, p = o->parent[i] + 1;
    sv tmp;
    memcpy(tmp, z[b], sizeof(sv));
    if (o->jtype[i] == 1) {
      Y[i] = force[6 + o->dof[i]] - svdot(o->S[i], z[b]);
      double s = o->invD[i] * Y[i];
      for (int k = 0; k < 6; k++) tmp[k] += o->h[i][k] * s;
    }
    sv pp;
    xf_force_inv(pp, o->Rp[b], o->rvec[i], tmp);
    for (int k = 0; k < 6; k++) z[p][k] += pp[k];
  }
  sv res;
  solve6(o->IA0, z[0], res);
  for (int k = 0; k < 6; k++) a[0][k] = o->fixed_base ? 0.0 : -res[k];
  for (int i = 0; i < n; i++) {
    int b = i + 1, p = o->parent[i] + 1;
    xf_motion(a[b], o->Rp[b], o->rvec[i], a[p]);
    if (o->jtype[i] == 1) {
      double qdd = o->invD[i] * (Y[i] - svdot(a[b], o->h[i]));
      out[6 + o->dof[i]] = qdd;
      for (int k = 0; k < 6; k++) a[b][k] += o->S[i][k] * qdd;
    }
  }
  m3tmulv(out, o->Rw[0], a[0]);
  m3tmulv(out + 3, o->Rw[0], a[0] + 3);
}

/* contact Jacobian (btMultiBody::fillConstraintJacobianMultiDof, linear direction only):
 * J = dir^T * d(point velocity)/d(u) for a point P (world) fixed to link `link` (-1 = base). */
static void point_jacobian(const trex_oracle* o, int link, const v3 P, const v3 dir, double* J) {
  memset(J, 0, sizeof(double) * o->nu);
  v3 rel = {P[0] - o->pos[0], P[1] - o->pos[1], P[2] - o->pos[2]}, t;
  v3cross(t, rel, dir);
  J[0] = t[0]; J[1] = t[1]; J[2] = t[2];
  J[3] = dir[0]; J[4] = dir[1]; J[5] = dir[2];
  for (int k = link; k >= 0; k = o->parent[k]) {
    if (o->jtype[k] != 1) continue;
    v3 aw, dw, piv, arm, c;
    m3tmulv(aw, o->Rw[k + 1], o->axis[k]);
    m3tmulv(dw, o->Rw[k + 1], o->dvec[k]);
    v3set(piv, o->comw[k + 1][0] - dw[0], o->comw[k + 1][1] - dw[1], o->comw[k + 1][2] - dw[2]);
    v3set(arm, P[0] - piv[0], P[1] - piv[1], P[2] - piv[2]);
    v3cross(c, aw, arm);
    J[6 + o->dof[k]] = v3dot(dir, c);
  }
}

static void candidate_world(const trex_oracle* o, int k, v3 P) {
  int b = o->cand_link[k] + 1;
  v3 t;
  m3tmulv(t, o->Rw[b], o->cand_local[k]);
  /* a sphere candidate touches the floor with its lowest point: centre - r * (0,0,1), a world offset */
  v3set(P, o->comw[b][0] + t[0], o->comw[b][1] + t[1], o->comw[b][2] + t[2] - o->cand_r[k]);
}

/* btMultiBodyConstraint::fillMultiBodyConstraint tail: unit-impulse response, jacDiagABInv, J.u */
static double finish_row(const trex_oracle* o, row_t* r, const double* u) {
  accel_deltas(o, r->J, r->dV);
  double d = 0, rel = 0;
  for (int k = 0; k < o->nu; k++) { d += r->J[k] * r->dV[k]; rel += r->J[k] * u[k]; }
  r->jdi = (d > 2.220446049250313e-16) ? 1.0 / d : 0.0;
  r->lam = 0;
  return rel;
}

/* btMultiBodyConstraintSolver::resolveSingleConstraintRowGeneric */
static double resolve_row(const trex_oracle* o, row_t* c, double* dv) {
  double delta = c->rhs; /* cfm = 0 */
  double jdv = 0;
  for (int k = 0; k < o->nu; k++) jdv += c->J[k] * dv[k];
  delta -= jdv * c->jdi;
  double sum = c->lam + delta;
  if (sum < c->lo) { delta = c->lo - c->lam; c->lam = c->lo; }
  else if (sum > c->hi) { delta = c->hi - c->lam; c->lam = c->hi; }
  else c->lam = sum;
  for (int k = 0; k < o->nu; k++) dv[k] += c->dV[k] * delta;
  return c->jdi != 0.0 ? delta / c->jdi : 0.0;
}

/* btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows */
static double resolve_cone(const trex_oracle* o, row_t* cA, row_t* cB, double* dv) {
  double jA = 0, jB = 0;
  for (int k = 0; k < o->nu; k++) { jA += cA->J[k] * dv[k]; jB += cB->J[k] * dv[k]; }
  double dB = cB->rhs - jB * cB->jdi, sumB = cB->lam + dB;
  double dA = cA->rhs - jA * cA->jdi, sumA = cA->lam + dA;
  double ang = atan2(sumA, sumB);
  double clipA = fabs(cA->lo * sin(ang)), clipB = fabs(cB->lo * cos(ang));
  if (sumA < -clipA) { dA = -clipA - cA->lam; cA->lam = -clipA; }
  else if (sumA > clipA) { dA = clipA - cA->lam; cA->lam = clipA; }
  else cA->lam = sumA;
  if (sumB < -clipB) { dB = -clipB - cB->lam; cB->lam = -clipB; }
  else if (sumB > clipB) { dB = clipB - cB->lam; cB->lam = clipB; }
  else cB->lam = sumB;
  for (int k = 0; k < o->nu; k++) dv[k] += cA->dV[k] * dA;
  for (int k = 0; k < o->nu; k++) dv[k] += cB->dV[k] * dB;
  double r = 0;
  if (cA->jdi != 0.0) r += dA / cA->jdi;
  if (cB->jdi != 0.0) r += dB / cB->jdi;
  return r;
}

/* ------------------------------------------------------------------ one stepSimulation
 * SURVEY.md Appendix A.3 steps 0-9. */
void trex_oracle_substep(trex_oracle* o, const double* target, double max_impulse) {
  const double dt = o->P[P_TIME_STEP] / o->n_sub;
  const int iters = (int)(o->P[P_SOLVER_ITERS] / o->n_sub);
  const int nu = o->nu, nd = o->ndof;
  double jt[MAXDOF] = {0}, acc[MAXU], u[MAXU], dv[MAXU];

  /* 0: pybullet joint damping as explicit joint torque (PhysicsServerCommandProcessor) */
  for (int k = 0; k < nd; k++) jt[k] = -o->damping[o->dof_link[k]] * o->qd[k];

  /* 1-2: kinematics at the current pose (contact detection uses these) */
  forward_kinematics(o);

  /* 3-4: gravity + ABA, velocities += dt * acc (clamped) */
  aba(o, jt, acc);
  apply_delta_vee(o, acc, dt);
  for (int k = 0; k < 3; k++) { u[k] = o->omega[k]; u[3 + k] = o->vel[k]; }
  for (int k = 0; k < nd; k++) u[6 + k] = o->qd[k];

  /* 5: constraint rows */
  row_t* rows = o->rows;
  int n_nc = 0;
  int n_limits = 0;
  for (int s = 0; s < o->n_order; s++) {
    int id = o->nc_order[s];
    if (id < nd) {
      /* btMultiBodyJointLimitConstraint: row 0 lower (dir +1), row 1 upper (dir -1), only when violated */
      int k = id, link = o->dof_link[k];
      for (int side = 0; side < 2; side++) {
        double pen = side == 0 ? o->q[k] - o->lower[link] : o->upper[link] - o->q[k];
        if (pen > 0) continue;
        row_t* r = &rows[n_nc++];
        memset(r->J, 0, sizeof(double) * nu);
        r->J[6 + k] = side == 0 ? 1.0 : -1.0;
        r->motor_dof = -1; r->cand = -1; r->fric_index = -1; r->limit_dof = k;
        double rel = finish_row(o, r, u);
        r->lo = 0; r->hi = o->P[P_LIMIT_MAX_IMPULSE];
        /* btMultiBodyJointLimitConstraint::createConstraintRows [RECALL]: with split impulse on (the default),
         * `penetration > m_splitImpulsePenetrationThreshold` (a SHALLOW violation, -0.04 < pen <= 0) combines the
         * positional (erp = m_erp) and the velocity term into m_rhs; a DEEP violation splits them and the positional
         * part goes to m_rhsPenetration, which the multibody solver never consumes: velocity part only. */
        double poserr = 0, velerr = -rel;
        if (pen > o->P[P_SPLIT_THRESH]) {
          poserr = -pen * o->P[P_ERP] / dt;
          r->rhs = (poserr + velerr) * r->jdi;
        } else {
          r->rhs = velerr * r->jdi;
        }
        n_limits++;
      }
    } else {
      /* btMultiBodyJointMotor: rhs velocity target = kp*(target-q)/dt + qd + kd*(0-qd) */
      int k = id - nd;
      row_t* r = &rows[n_nc++];
      memset(r->J, 0, sizeof(double) * nu);
      r->J[6 + k] = 1.0;
      r->motor_dof = k; r->cand = -1; r->fric_index = -1; r->limit_dof = -1;
      double rel = finish_row(o, r, u);
      double qd = o->qd[k];
      double vt = o->P[P_KP] * (target[k] - o->q[k]) / dt + qd + o->P[P_KD] * (0.0 - qd);
      r->rhs = (vt - rel) * r->jdi;
      r->lo = -max_impulse; r->hi = max_impulse;
    }
  }
  /* contacts: candidate points against the floor plane, normal (0,0,1), btPlaneSpace1 tangents */
  uint32_t sig_k[2] = {0, 0}; /* active-set signature: the candidates with rows */
  int n_normal = 0;
  row_t* nrm = rows + n_nc;
  const v3 nz = {0, 0, 1}, t1 = {0, -1, 0}, t2 = {1, 0, 0};
  if (o->contacts_on) {
    /* contact capacity (definition N2, shared with the kernels): at most max_contacts points have rows; when more
     * candidates are inside the breaking distance the deepest ones are kept (ties by candidate index) */
    int keep[MAXC];
    double zc[MAXC];
    int n_in = 0;
    for (int c = 0; c < o->ncand; c++) {
      v3 P;
      candidate_world(o, c, P);
      zc[c] = P[2];
      keep[c] = (P[2] - o->P[P_FLOOR]) < o->P[P_BREAKING];
      n_in += keep[c];
    }
    const int cap = (int)o->P[P_MAX_CONTACTS];
    if (n_in > cap) {
      for (int c = 0; c < o->ncand; c++) {
        if (!keep[c]) continue;
        int better = 0;
        for (int c2 = 0; c2 < o->ncand; c2++)
          if (c2 != c && (zc[c2] - o->P[P_FLOOR]) < o->P[P_BREAKING] && (zc[c2] < zc[c] || (zc[c2] == zc[c] && c2 < c))) better++;
        if (better >= cap) keep[c] = 0;
      }
    }
    for (int c = 0; c < o->ncand; c++) {
      v3 P;
      candidate_world(o, c, P);
      double dist = P[2] - o->P[P_FLOOR];
      if (keep[c]) { if (c < 32) sig_k[0] |= 1u << c; else sig_k[1] |= 1u << (c - 32); }
      if (!keep[c]) { o->lam_cache[c] = 0; continue; }
      row_t* r = &nrm[n_normal++];
      point_jacobian(o, o->cand_link[c], P, nz, r->J);
      r->cand = c; r->motor_dof = -1; r->fric_index = -1; r->limit_dof = -1;
      double rel = finish_row(o, r, u);
      double pen = dist + o->P[P_LINEAR_SLOP];
      double poserr = 0, velerr = -rel; /* restitution 0 */
      if (pen > 0) velerr -= pen / dt;
      else poserr = -pen * o->P[P_CONTACT_ERP] / dt;
      r->rhs = (poserr + velerr) * r->jdi;
      r->lo = 0; r->hi = 1e10;
      r->mu = o->P[P_FRICTION];
    }
  }
  row_t* fr = nrm + n_normal;
  for (int c = 0; c < n_normal; c++) {
    v3 P;
    candidate_world(o, nrm[c].cand, P);
    for (int d = 0; d < 2; d++) {
      row_t* r = &fr[2 * c + d];
      point_jacobian(o, o->cand_link[nrm[c].cand], P, d == 0 ? t1 : t2, r->J);
      r->cand = nrm[c].cand; r->motor_dof = -1; r->fric_index = c; r->limit_dof = -1;
      double rel = finish_row(o, r, u);
      r->rhs = -rel * r->jdi;
      r->mu = o->P[P_FRICTION];
      r->lo = 0; r->hi = 0;
    }
  }
  memset(dv, 0, sizeof(double) * nu);
  /* warm start of the normal rows (cached impulse * warmstartingFactor) */
  for (int c = 0; c < n_normal; c++) {
    double imp = o->lam_cache[nrm[c].cand] * o->P[P_WARMSTART];
    nrm[c].lam = imp;
    if (imp != 0.0)
      for (int k = 0; k < nu; k++) dv[k] += nrm[c].dV[k] * imp;
  }

  /* 6: PGS (btSequentialImpulseConstraintSolver::solveGroupCacheFriendlyIterations +
   *          btMultiBodyConstraintSolver::solveSingleIteration) */
  int it_done = 0;
  for (int it = 0; it < iters; it++) {
    double resid = 0;
    for (int j = 0; j < n_nc; j++) {
      int idx = (it & 1) ? j : n_nc - 1 - j;
      double r = resolve_row(o, &rows[idx], dv);
      if (r * r > resid) resid = r * r;
    }
    for (int c = 0; c < n_normal; c++) {
      double r = resolve_row(o, &nrm[c], dv);
      if (r * r > resid) resid = r * r;
    }
    for (int c = 0; c < n_normal; c++) {
      double tot = nrm[c].lam;
      row_t *a = &fr[2 * c], *b = &fr[2 * c + 1];
      a->lo = -(a->mu * tot); a->hi = a->mu * tot;
      b->lo = -(b->mu * tot); b->hi = b->mu * tot;
      double r = resolve_cone(o, a, b, dv);
      if (r * r > resid) resid = r * r;
    }
    if (it < 512) o->resid_hist[it] = resid;
    it_done = it + 1;
    if (resid <= o->P[P_RESIDUAL] || it >= iters - 1) break;
  }

  { /* active-set signature of the step, folded per substep exactly like the kernels do (trex_core.h: StepStats):
     * K (candidates with rows), then L (limit rows that ended with a positive impulse), M (motors that ended on their bound),
     * N | F << 16 (contact slots with a positive normal impulse / with the friction pair on the cone), iterations */
    uint32_t L = 0, M = 0, N = 0, F = 0, h = o->sig;
    for (int j = 0; j < n_nc; j++) {
      if (rows[j].limit_dof >= 0 && rows[j].lam > 0) L |= 1u << rows[j].limit_dof;
      if (rows[j].motor_dof >= 0 && fabs(rows[j].lam) >= max_impulse) M |= 1u << rows[j].motor_dof;
    }
    for (int c = 0; c < n_normal; c++) {
      if (!(nrm[c].lam > 0)) continue;
      N |= 1u << c;
      double lim = nrm[c].mu * nrm[c].lam, a = fr[2 * c].lam, b = fr[2 * c + 1].lam;
      if (a * a + b * b >= 0.9999 * (lim * lim)) F |= 1u << c;
    }
#define SIG_MIX(h, w) (((h) ^ (uint32_t)(w)) * 16777619u)
    h = SIG_MIX(SIG_MIX(h, sig_k[0]), sig_k[1]) & 0xffffffu;
    h = SIG_MIX(SIG_MIX(SIG_MIX(SIG_MIX(h, L), M), N | (F << 16)), (uint32_t)it_done) & 0xffffffu;
#undef SIG_MIX
    o->sig = h;
    o->sig_words[0] = sig_k[0]; o->sig_words[1] = sig_k[1]; o->sig_words[2] = L; o->sig_words[3] = M;
    o->sig_words[4] = N | (F << 16); o->sig_words[5] = (uint32_t)it_done;
  }

  /* 7: velocities += dv (clamped); write back impulses */
  apply_delta_vee(o, dv, 1.0);
  for (int k = 0; k < nd; k++) o->tau[k] = 0;
  for (int j = 0; j < n_nc; j++)
    if (rows[j].motor_dof >= 0) o->tau[rows[j].motor_dof] = rows[j].lam / dt;
  for (int c = 0; c < n_normal; c++) o->lam_cache[nrm[c].cand] = nrm[c].lam;

  /* 8: btMultiBody::stepPositionsMultiDof with the NEW velocities */
  for (int k = 0; k < 3; k++) o->pos[k] += dt * o->vel[k];
  {
    double fAngle = v3norm(o->omega);
    const double thresh = 0.5 * 1.5707963267948966;
    if (fAngle * dt > thresh) fAngle = 0.5 * 1.5707963267948966 / dt;
    double sc;
    if (fAngle < 0.001) sc = 0.5 * dt - (dt * dt * dt) * 0.020833333333 * fAngle * fAngle;
    else sc = sin(0.5 * fAngle * dt) / fAngle;
    double ax = o->omega[0] * sc, ay = o->omega[1] * sc, az = o->omega[2] * sc, aw = cos(fAngle * dt * 0.5);
    /* base->world quaternion: q <- dq * q  (Bullet: world->base quat <- quat * conj(dq)) */
    double x = o->quat[0], y = o->quat[1], z = o->quat[2], w = o->quat[3];
    double nx = aw * x + ax * w + ay * z - az * y;
    double ny = aw * y - ax * z + ay * w + az * x;
    double nz_ = aw * z + ax * y - ay * x + az * w;
    double nw = aw * w - ax * x - ay * y - az * z;
    double inv = 1.0 / sqrt(nx * nx + ny * ny + nz_ * nz_ + nw * nw);
    o->quat[0] = nx * inv; o->quat[1] = ny * inv; o->quat[2] = nz_ * inv; o->quat[3] = nw * inv;
  }
  for (int k = 0; k < nd; k++) o->q[k] += dt * o->qd[k];

  o->last_iters = it_done;
  o->last_contacts = n_normal;
  o->last_limits = n_limits;
  o->total_iters += it_done;
  o->total_substeps += 1;
}

/* ------------------------------------------------------------------ env surface */
void trex_oracle_head_position(trex_oracle* o, double* xyz) {
  forward_kinematics(o);
  v3cpy(xyz, o->comw[o->head_link + 1]); /* getLinkState()[0] = link COM, world (trex_robot.py:330-335) */
}

static void observe(trex_oracle* o, double* obs) { /* trex_robot.py:359-365 */
  if (!obs) return;
  for (int k = 0; k < 25; k++) {
    int d = o->obs_dof[k];
    obs[k] = o->q[d]; obs[25 + k] = o->qd[d]; obs[50 + k] = o->tau[d];
  }
}

static double reward(trex_oracle* o) { /* trex_env.py:186-196 ; trex_robot.py:367-375 */
  double p[3];
  trex_oracle_head_position(o, p);
  double power = 0;
  for (int k = 0; k < 25; k++) { int d = o->obs_dof[k]; power += fabs(o->qd[d] * o->tau[d]); }
  double th = o->P[P_TARGET_HEIGHT];
  double station = o->w_drift * (p[0] * p[0] + p[1] * p[1]);
  double lifting = o->w_dist * ((th - p[2]) * (th - p[2]));
  double energy = o->w_energy * power;
  o->rterm[0] = lifting; o->rterm[1] = station; o->rterm[2] = energy;
  return -lifting - station - energy;
}

void trex_oracle_reset(trex_oracle* o, double* obs) {
  /* trex_robot.py:57-65,300-309: base COM frame at [0,0,reset_height], identity, zero velocity,
   * joints zero + starting configuration, all motors zero gain / zero force */
  v3set(o->pos, 0, 0, o->P[P_RESET_HEIGHT]);
  o->quat[0] = o->quat[1] = o->quat[2] = 0; o->quat[3] = 1;
  v3set(o->omega, 0, 0, 0); v3set(o->vel, 0, 0, 0);
  for (int k = 0; k < o->ndof; k++) { o->q[k] = o->start_q[o->dof_link[k]]; o->qd[k] = 0; o->tau[k] = 0; }
  memset(o->lam_cache, 0, sizeof(o->lam_cache));
  double zero[MAXDOF] = {0};
  /* trex_env.py:120: one stepSimulation inside reset. Motor gains kp=kd=0, force 0:
   * target velocity = qd + 0, bounds +-0 -> rows exist but carry no impulse. */
  double kp = o->P[P_KP], kd = o->P[P_KD];
  o->P[P_KP] = 0; o->P[P_KD] = 0;
  trex_oracle_substep(o, zero, 0.0);
  o->P[P_KP] = kp; o->P[P_KD] = kd;
  observe(o, obs);
}

void trex_oracle_step(trex_oracle* o, const double* action, double* obs, double* rew) {
  const double dt = o->P[P_TIME_STEP] / o->n_sub;
  double target[MAXDOF];
  for (int k = 0; k < 25; k++) { /* np.clip(action, low, high)  trex_env.py:147 */
    int d = o->obs_dof[k], link = o->dof_link[d];
    target[d] = clampd(action[k], o->lower[link], o->upper[link]);
  }
  o->sig = 0;
  for (int s = 0; s < o->n_sub; s++) /* trex_env.py:148-150 ; max impulse = force*dt */
    trex_oracle_substep(o, target, o->P[P_MAX_TORQUE] * dt);
  observe(o, obs);
  double r = reward(o);
  if (rew) *rew = r;
}

void trex_oracle_run(trex_oracle* o, const double* actions, int T, double* obs_out, double* reward_out) {
  for (int t = 0; t < T; t++)
    trex_oracle_step(o, actions + 25 * (size_t)t, obs_out ? obs_out + 75 * (size_t)t : NULL, reward_out ? reward_out + t : NULL);
}

void trex_oracle_momentum(trex_oracle* o, double* out) {
  forward_kinematics(o);
  /* link velocities in link frames, as in aba() pass 1 */
  static __thread sv vel[MAXL + 1];
  m3mulv(vel[0], o->Rw[0], o->omega);
  m3mulv(vel[0] + 3, o->Rw[0], o->vel);
  double P[3] = {0, 0, 0}, L[3] = {0, 0, 0}, ke = 0, mt = 0, com[3] = {0, 0, 0};
  for (int b = 0; b <= o->n; b++) {
    if (b > 0) {
      int i = b - 1;
      xf_motion(vel[b], o->Rp[b], o->rvec[i], vel[o->parent[i] + 1]);
      if (o->jtype[i] == 1)
        for (int k = 0; k < 6; k++) vel[b][k] += o->S[i][k] * o->qd[o->dof[i]];
    }
    double m = o->mass[b];
    v3 vw, Iw = {o->inertia[b][0] * vel[b][0], o->inertia[b][1] * vel[b][1], o->inertia[b][2] * vel[b][2]}, Lw, rxp;
    m3tmulv(vw, o->Rw[b], vel[b] + 3);
    m3tmulv(Lw, o->Rw[b], Iw);
    v3 mv = {m * vw[0], m * vw[1], m * vw[2]};
    v3cross(rxp, o->comw[b], mv);
    for (int k = 0; k < 3; k++) { P[k] += mv[k]; L[k] += Lw[k] + rxp[k]; com[k] += m * o->comw[b][k]; }
    ke += 0.5 * m * v3dot(vw, vw) + 0.5 * v3dot(vel[b], Iw);
    mt += m;
  }
  memcpy(out, P, 24); memcpy(out + 3, L, 24);
  out[6] = ke; out[7] = mt;
  out[8] = com[0] / mt; out[9] = com[1] / mt; out[10] = com[2] / mt;
}

void trex_oracle_minv_column(trex_oracle* o, int dofidx, double* out) {
  double jt[MAXDOF] = {0}, acc[MAXU], f[MAXU] = {0};
  forward_kinematics(o);
  /* aba() caches h/invD/IA0 for the current pose; velocities do not enter those */
  aba(o, jt, acc);
  f[dofidx] = 1.0;
  accel_deltas(o, f, out);
}

void trex_oracle_residual_history(const trex_oracle* o, double* out, int n) {
  for (int i = 0; i < n && i < 512; i++) out[i] = i < o->last_iters ? o->resid_hist[i] : 0.0;
}

void trex_oracle_candidate_position(trex_oracle* o, int k, double* xyz) {
  forward_kinematics(o);
  candidate_world(o, k, xyz);
}
