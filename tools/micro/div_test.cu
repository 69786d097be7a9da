// div_fast / hypot_fast / tanh_fast of jaero_b200/csrc/demod_device.cuh against the exactly rounded / library results
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "../../jaero_b200/csrc/demod_device.cuh"
__global__ void k(const double *a, const double *b, double *q, double *h, double *t, double *q0, double *h0, double *t0, int n)
{ int i = blockIdx.x*blockDim.x+threadIdx.x; if (i<n) { q[i] = jb::div_fast(a[i], b[i]); q0[i] = a[i] / b[i]; h[i] = jb::hypot_fast(a[i], b[i]); h0[i] = hypot(a[i], b[i]); t[i] = jb::tanh_fast(a[i]); t0[i] = tanh(a[i]); } }
int main(){ const int n=1<<22; double *a,*b,*q,*h,*t,*q0,*h0,*t0; for (double **p : {&a,&b,&q,&h,&t,&q0,&h0,&t0}) cudaMallocManaged(p,n*8);
 srand(2); for(int i=0;i<n;i++){ a[i]=(rand()/(double)RAND_MAX*2-1)*4.0; b[i]=1e-6+rand()/(double)RAND_MAX*3.0; if (i%5==0) { a[i]=1.414213562; b[i]=0.05+rand()/(double)RAND_MAX; } }
 k<<<(n+255)/256,256>>>(a,b,q,h,t,q0,h0,t0,n); if (cudaDeviceSynchronize()!=cudaSuccess){printf("cuda error\n");return 1;}
 int dq=0,dh=0,dt=0; double mh=0,mt=0; for(int i=0;i<n;i++){ if(q[i]!=q0[i])dq++; if(h[i]!=h0[i]){dh++; double u=fabs(h[i]-h0[i])/(nextafter(h0[i],1e300)-h0[i]); if(u>mh)mh=u;} if(t[i]!=t0[i]){dt++; double u=fabs(t[i]-t0[i])/(nextafter(fabs(t0[i]),1e300)-fabs(t0[i])); if(u>mt)mt=u;} }
 printf("div_fast != IEEE division: %d of %d; hypot_fast != hypot: %d (max %.2f ulp); tanh_fast != tanh: %d (max %.2f ulp)\n", dq,n,dh,mh,dt,mt); return 0; }
