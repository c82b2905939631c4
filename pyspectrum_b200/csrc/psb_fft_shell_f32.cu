#include "psb_fft_shell.cuh"
namespace psb {
template int fft_shell_pair<float>(const Cx<float>*, const unsigned short*, int, int, int, int, int, Cx<float>*, Cx<float>*, float*, float*, double*, const float*, unsigned int*, int, const Cx<float>*, cudaStream_t, const long long*, int, int);
}
