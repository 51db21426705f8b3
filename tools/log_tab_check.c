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
int main(){
  for(int i=0;i<128;i++){ double c = 1.0 + (i+0.5)/128; double inv = 1.0/c; T[i][0]=inv; T[i][1] = -(double)logl((long double)inv); }
  double maxabs=0, maxrel=0; srand(1);
  for(long n=0;n<20000000;n++){
    double x;
    if(n&1){ x = exp(-30.0*rand()/RAND_MAX); } else { x = 1.0 - 1e-3*rand()/RAND_MAX*rand()/RAND_MAX; }
    long double ref = logl((long double)x);
    double got = flog(x);
    double ea = fabs((double)(got-ref));
    double er = ref!=0 ? ea/fabs((double)ref) : 0;
    double elib = fabs((double)(log(x)-ref));
    if(ea>maxabs) maxabs=ea;
    if(fabs((double)ref)>1e-3 && er>maxrel) maxrel=er;
  }
  printf("max abs err %.3g  max rel err (|log|>1e-3) %.3g\n", maxabs, maxrel);
  return 0;
}
