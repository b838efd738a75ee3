// TEST-ONLY host emulation of the kernel source (csrc/scp_device.inl compiled as
// plain C++ with SCP_EMU: a phase = a sequential loop over thread ids).  It
// exists to debug the solver logic in a container without a GPU.  The product
// package never builds, loads or falls back to this file.
#define SCP_EMU 1
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../ba-path-planning_b200/csrc/scp_device.inl"
#include "../../ba-path-planning_b200/csrc/scp_tables.h"

extern "C" int scp_emu_solve_batch(const scp_b200_problem* prob, int B, const double* p0, const double* v0,
                                   const double* pf, const double* vf, double* acc, double* pos,
                                   double* vel, scp_b200_record* rec, int nthreads) {
  const int resumable = std::getenv("SCP_EMU_WHOLE") ? 0 : 1;
  using namespace scp;
  const int N = prob->n_agents, K = prob->n_steps;
  HostTables ht = build_host_tables(*prob);
  Params g;
  g.pb = *prob;
  g.L = make_layout(N, K);
  const double* tb = ht.blob.data();
  g.tb.B1 = tb + ht.oB1; g.tb.B2 = tb + ht.oB2; g.tb.rj = tb + ht.orj; g.tb.ra = tb + ht.ora;
  g.tb.rv = tb + ht.orv; g.tb.rp = tb + ht.orp; g.tb.rc = tb + ht.orc;
  std::vector<double> wd(g.L.n_double, std::nan(""));   // poisoned: the GPU scratch is uninitialised too
  std::vector<int> wi(g.L.n_int);
  { unsigned st = 12345u; for (auto& v : wi) { st = st * 1664525u + 1013904223u; v = (int)(st >> 4) - (1 << 26); } }   // garbage like GPU scratch
  std::vector<double> sm(sh_doubles(RED) + (size_t)K * K, 0.0);
  for (int b = 0; b < B; ++b) {
    Ctx c;
    c.nthreads = nthreads; c.N = N; c.K = K; c.Q = 2 * N; c.g = &g;
    c.team = 1; c.tid0 = 0; c.np = nthreads < 512 ? nthreads : 512; c.rs = RED; c.sh = sm.data();
    c.wd = wd.data(); c.wi = wi.data(); c.sm = sm.data(); c.nmat = nullptr; c.nmat_in_smem = 1; c.fused_epl = 0; c.fused_rows = nullptr;
    c.a_x = c.wd + g.L.x; c.a_rhs = c.wd + g.L.rhs; c.a_vj = c.wd + g.L.vj; c.a_va = c.wd + g.L.va;
    c.a_vv = c.wd + g.L.vv; c.a_vp = c.wd + g.L.vp; c.a_P = c.wd + g.L.P; c.a_F = c.wd + g.L.F;
    c.pol_smem = nullptr; c.pol_smem_doubles = 0;
    for (int a = 0; a < 7; ++a) { c.hot_s[a] = nullptr; c.hot_g[a] = nullptr; }
    size_t s2 = (size_t)b * N * 2, s3 = (size_t)b * N * K * 2;
    c.p0 = p0 + s2; c.v0 = v0 + s2; c.pf = pf + s2; c.vf = vf + s2;
    c.acc = acc + s3; c.pos = pos + s3; c.vel = vel + s3; c.rec = rec + b;
    int mode = resumable ? 2 : 0;
    while (!solve_scenario(c, mode)) mode = 1;   // quanta until the scenario is finished
  }
  return 0;
}

extern "C" void scp_emu_default_problem(scp_b200_problem* p, int n, double T, double h, double R);
#include "../../ba-path-planning_b200/csrc/scp_defaults.h"
extern "C" void scp_emu_default_problem(scp_b200_problem* p, int n, double T, double h, double R) {
  scp_fill_default_problem(p, n, T, h, R);
}
