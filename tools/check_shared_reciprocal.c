// Evidence for xf_div_shared (xfluids_b200/csrc/xf_math.cuh): a/b from the shared correctly rounded reciprocal of b, one residual fma and
// one correction fma, equals the IEEE quotient.  gcc -O2 -ffp-contract=off -mfma tools/check_shared_reciprocal.c -lm && ./a.out 300000000
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
static uint64_t s=88172645463325252ULL;
static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
static inline double urand(){return (rnd()>>11)*(1.0/9007199254740992.0);}
int main(int argc,char**argv){
  long n=atol(argv[1]); long bad=0;
  for(long i=0;i<n;i++){
    double b = 200.0 + urand()*6000.0;          // T range
    if(i&1){ uint64_t m=rnd(); union{uint64_t u;double d;}x; x.u=(m&0x000fffffffffffffULL)|0x4090000000000000ULL; b=x.d; } // random mantissa, 1024..2048
    double a = (urand()-0.5)*pow(10.0,(double)(rnd()%16)-4.0); // coefficient magnitudes 1e-4..1e11
    double r = 1.0/b;
    double q0 = a*r;
    double rem = fma(-q0,b,a);
    double q = fma(rem,r,q0);
    double t = a/b;
    if(q!=t){bad++; if(bad<5) printf("a=%a b=%a q=%a t=%a\n",a,b,q,t);}
  }
  printf("n=%ld mismatches=%ld\n",n,bad); return 0;}
