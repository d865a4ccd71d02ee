// msv_wave_kernels.cuh -- single-sequence wavefront kernel (placeholder; filled in below).
#pragma once
