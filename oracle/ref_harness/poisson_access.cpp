// oracle/ref_harness/poisson_access.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference keeps the potential in a file-static vector (/root/reference/src/poisson.cpp:9).
// This translation unit compiles the UNMODIFIED reference source by inclusion (it is not copied
// into the repo; REF_SRC_POISSON is passed by oracle/Makefile) and adds one accessor so the
// oracle can report phi, one of the graded fields.
#include REF_SRC_POISSON

namespace poisson {
const std::vector<double>& oracle_phi() { return phi; }
}
