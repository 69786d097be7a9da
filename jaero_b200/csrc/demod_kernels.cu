// Single translation unit for the two segment kernels (they share the __constant__ tap table).
// Built with -fmad=false: see demod_device.cuh.
#include "oqpsk_demod.cu"
#include "oqpsk_pipe.cu"
#include "msk_demod.cu"
#include "msk_pipe.cu"
