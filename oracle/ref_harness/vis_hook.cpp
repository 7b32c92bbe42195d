// oracle/ref_harness/vis_hook.cpp -- TEST INFRASTRUCTURE ONLY.
// Replaces /root/reference/src/visualize.cpp (needs OpenCV, absent here) with the three entry
// points LBmethod::Run_simulation calls (/root/reference/src/plasma.cpp:465,516-522,525;
// declared in /root/reference/include/visualize.hpp:50-65).  UpdateVisualization receives the
// 15 host fields of step t -- pre-collision moments and the field after that step's Poisson
// solve -- and writes them, plus phi, as raw float64 in the order of its parameter list.
#include "plasma.hpp"
#include "vis_hook.hpp"

#include <cstdio>
#include <stdexcept>

namespace poisson { const std::vector<double>& oracle_phi(); }

namespace ref_hook {
Config& config() { static Config c; return c; }
}

namespace visualize {

void InitVisualization(const int, const int, const int) {}
void CloseVisualization() {}

void UpdateVisualization(const int t, const int NX, const int NY,
    const std::vector<double>& ux_e,  const std::vector<double>& uy_e,
    const std::vector<double>& ux_i,  const std::vector<double>& uy_i,
    const std::vector<double>& ux_n,  const std::vector<double>& uy_n,
    const std::vector<double>& T_e,   const std::vector<double>& T_i,
    const std::vector<double>& T_n,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i,
    const std::vector<double>& rho_n, const std::vector<double>& rho_q,
    const std::vector<double>& Ex,    const std::vector<double>& Ey)
{
    const ref_hook::Config& c = ref_hook::config();
    if (c.out_dir.empty()) return;
    if (!c.dump_all && !c.dump_steps.count(t)) return;
    char name[512];
    std::snprintf(name, sizeof(name), "%s/fields_t%05d.f64", c.out_dir.c_str(), t);
    FILE* fp = std::fopen(name, "wb");
    if (!fp) throw std::runtime_error(std::string("cannot open ") + name);
    const size_t n = static_cast<size_t>(NX) * NY;
    const std::vector<double>* fields[15] = { &ux_e, &uy_e, &ux_i, &uy_i, &ux_n, &uy_n, &T_e, &T_i, &T_n,
                                              &rho_e, &rho_i, &rho_n, &rho_q, &Ex, &Ey };
    for (const auto* f : fields) std::fwrite(f->data(), sizeof(double), n, fp);
    std::vector<double> phi = poisson::oracle_phi();
    phi.resize(n, 0.0);
    std::fwrite(phi.data(), sizeof(double), n, fp);
    std::fclose(fp);
}

} // namespace visualize
