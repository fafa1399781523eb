// Process-wide count of kernels launched by this library (reported as bench.py's gpu_launches).
#pragma once
#include <stdint.h>
namespace dsee {
void count_launch(int n = 1);
int64_t launch_count();
}  // namespace dsee
