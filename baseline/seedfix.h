// Force-included (-include) when compiling the reference's rchol_lap.cpp ONLY.
// The reference seeds each thread_local std::mt19937 from std::random_device
// (/root/reference/c++/rchol_lap/rchol_lap.cpp:685), which makes the factor
// non-reproducible.  Without touching the reference sources we substitute a
// deterministic "device" whose value the producer shim sets before each call.
#pragma once
#include <random>
extern "C" unsigned rchol_b200_fixed_seed;
namespace std {
struct rchol_b200_fixed_rd {
  typedef unsigned result_type;
  unsigned operator()() { return rchol_b200_fixed_seed; }
};
}  // namespace std
#define random_device rchol_b200_fixed_rd
