// Accuracy of log_tab() (basevar_b200/csrc/bv_em_kernels.cuh) against 80-bit logl: the same table, the same operations, on the host.
//   gcc -O2 -o /tmp/log_tab_check tools/log_tab_check.c -lm && /tmp/log_tab_check [n]
// (-ffp-contract=off is not needed: the fused operations are spelled fma() here as on the device)
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static double T[128][2];
static double flog(double x){
  uint64_t u; memcpy(&u,&x,8);
  uint32_t hi = u>>32;
  int e = (int)(hi>>20) - 1023;
  int i = (hi>>13)&0x7f;
  uint64_t mu = (u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
  double z; memcpy(&z,&mu,8);
  double r = fma(z, T[i][0], -1.0);
  double p = fma(r, -1.0/6, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 1.0/3);
  p = fma(r, p, -0.5);
  p = fma(r*r, p, r);           // r + r^2 * (...)
  double base = fma((double)e, 0.693147180559945309417, T[i][1]);
  return base + p;
}
int main(int argc, char** argv){
  long N = argc > 1 ? atol(argv[1]) : 20000000;
  for(int i=0;i<128;i++){ double c = 1.0 + (i+0.5)/128; double inv = 1.0/c; T[i][0]=inv; T[i][1] = -(double)logl((long double)inv); }
  double maxabs=0, maxrel=0, maxulp=0, maxabs_near1=0, maxscaled=0; srand(1);
  for(long n=0;n<N;n++){
    double x;
    if(n&1){ x = exp(-30.0*rand()/RAND_MAX); } else { x = 1.0 - 1e-3*rand()/RAND_MAX*rand()/RAND_MAX; }
    long double ref = logl((long double)x);
    double got = flog(x);
    double ea = fabs((double)(got-ref));
    double er = ref!=0 ? ea/fabs((double)ref) : 0;
    double elib = fabs((double)(log(x)-ref));
    if(ea>maxabs) maxabs=ea;
    { double sc = fabs((double)ref) > 0.693147180559945 ? fabs((double)ref) : 0.693147180559945; if(ea/sc>maxscaled) maxscaled=ea/sc; }
    if(fabs((double)ref)<0.05 && ea>maxabs_near1) maxabs_near1=ea;
    if(fabs((double)ref)>=0.05){ int ex; frexp((double)ref,&ex); double u = ea/ldexp(1.0, ex-53); if(u>maxulp) maxulp=u; }
    if(fabs((double)ref)>1e-3 && er>maxrel) maxrel=er;
  }
  printf("max abs err %.3g  max rel err (|log|>1e-3) %.3g\n", maxabs, maxrel);
  printf("max_err_over_max_abs_log_ln2 %.4g\n", maxscaled);
  printf("max_ulp_err_where_abs_log_ge_0.05 %.3f\nmax_abs_err_where_abs_log_lt_0.05 %.3g\n", maxulp, maxabs_near1);
  return 0;
}
