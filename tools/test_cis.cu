#include <cstdio>
#include <cmath>
#include <vector>
#include "afr_common.cuh"
namespace afr { void set_error(const std::string&) {} int fail(const std::string&) {return 1;} }
__global__ void k(const double* p, double* o, int n) { int i = blockIdx.x*blockDim.x+threadIdx.x; if (i<n) { auto z = afr::cis_fast(p[i]); o[2*i]=z.re; o[2*i+1]=z.im; } }
int main() {
  int n = 1<<20; std::vector<double> p(n), o(2*n);
  srand(1);
  for (int i=0;i<n;i++){ double u = rand()/(double)RAND_MAX; double mag = pow(10.0, -8 + 14.2*(rand()/(double)RAND_MAX)); p[i] = (u-0.5)*2*mag; }
  p[0]=0; p[1]=1e6; p[2]=-1e6; p[3]=M_PI/4; p[4]=-M_PI/4; p[5]=M_PI/2; p[6]=1.5e6; p[7]=NAN;
  double *dp,*dout; cudaMalloc(&dp,n*8); cudaMalloc(&dout,2*n*8);
  cudaMemcpy(dp,p.data(),n*8,cudaMemcpyHostToDevice);
  k<<<(n+255)/256,256>>>(dp,dout,n); cudaMemcpy(o.data(),dout,2*n*8,cudaMemcpyDeviceToHost);
  double maxe=0; int arg=0;
  for(int i=0;i<n;i++){ if (std::isnan(p[i])) continue; double e = fmax(fabs(o[2*i]-cos(p[i])), fabs(o[2*i+1]-sin(p[i]))); if (e>maxe){maxe=e;arg=i;} }
  printf("max abs err %.3e at p=%.17g  (nan-> %f %f)\n", maxe, p[arg], o[14], o[15]);
  return 0;
}
