#include "psb_fft_shell.cuh"
namespace psb {
template int fft_shell_pair<double>(const Cx<float>*, const unsigned short*, int, int, int, int, int, Cx<double>*, Cx<double>*, double*, double*, double*, const float*, unsigned int*, int, const Cx<double>*, cudaStream_t, const long long*, int, int);
}
