// oracle/ref_harness/vis_hook.hpp -- TEST INFRASTRUCTURE ONLY.
#pragma once
#include <set>
#include <string>
namespace ref_hook {
struct Config {
    std::string out_dir;       // where fields_tNNNNN.f64 files go ("" = no dumps)
    std::set<int> dump_steps;  // time steps whose 15 fields + phi are written
    bool dump_all = false;
};
Config& config();
}
