// Force-included when compiling the reference's rchol_parallel.cpp ONLY: renames its call to
// find_separator(...) so that the producer shim can observe the partition sizes
// (`Separator_info::val`), which the reference computes (rchol_parallel.cpp:62-70) but never returns.
#pragma once
#define find_separator rchol_b200_find_separator_hook
