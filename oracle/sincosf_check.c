// TEST INFRASTRUCTURE — exhaustive check of the restated glibc sinf/cosf algorithm (the one
// box2d_rs_b200/csrc/b2g_math.h evaluates on the device) against the host libm, over all 2^32
// float inputs:  gcc -O2 -ffp-contract=off -mfma -DUSE_FMA sincosf_check.c -lm && ./a.out 0 0x100000000
// Result on glibc 2.39 / x86-64 (FMA ifunc variant): 0 mismatches for sin and cos.  Without
// -DUSE_FMA (the non-FMA build of glibc) 17 inputs per sign differ in the last bit.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
typedef struct { double sign[4]; double hpi_inv, hpi, c0,c1,c2,c3,c4,s1,s2,s3; } sincos_t;
static const sincos_t T[2] = {
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
static const uint32_t inv_pio4[24] = {0xa2,0xa2f9,0xa2f983,0xa2f9836e,0xf9836e4e,0x836e4e44,0x6e4e4415,0x4e441529,0x441529fc,0x1529fc27,0x29fc2757,0xfc2757d1,0x2757d1f5,0x57d1f534,0xd1f534dd,0xf534ddc0,0x34ddc0db,0xddc0db62,0xc0db6295,0xdb629599,0x6295993c,0x95993c43,0x993c4390,0x3c439041};
#ifdef USE_FMA
#define MA(a,b,c) fma((a),(b),(c))
#else
#define MA(a,b,c) ((a)*(b)+(c))
#endif
static inline uint32_t asuint(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static inline uint32_t abstop12(float x){return (asuint(x)>>20)&0x7ff;}
static inline float poly(double x, double x2, const sincos_t*p, int n){
  if((n&1)==0){ double x3=x*x2; double s1=MA(x2,p->s3,p->s2); double x7=x3*x2; double s=MA(x3,p->s1,x); return (float)MA(x7,s1,s);}
  else { double x4=x2*x2; double c2=MA(x2,p->c4,p->c3); double c1=MA(x2,p->c1,p->c0); double x6=x4*x2; double c=MA(x4,p->c2,c1); return (float)MA(x6,c2,c);}
}
static inline double reduce_fast(double x,const sincos_t*p,int*np){ double r=x*p->hpi_inv; int n=((int32_t)r+0x800000)>>24; *np=n; 
#ifdef USE_FMA
 return fma(-(double)n, p->hpi, x);
#else
 return x-n*p->hpi;
#endif
}
static inline double reduce_large(uint32_t xi,int*np){ const uint32_t*arr=&inv_pio4[(xi>>26)&15]; int shift=(xi>>23)&7; uint64_t n,res0,res1,res2; xi=(xi&0xffffff)|0x800000; xi<<=shift; res0=xi*arr[0]; res1=(uint64_t)xi*arr[4]; res2=(uint64_t)xi*arr[8]; res0=(res2>>32)|(res0<<32); res0+=res1; n=(res0+(1ULL<<61))>>62; res0-=n<<62; double x=(int64_t)res0; *np=n; return x*0x1.921FB54442D18p-62;}
float my_sinf(float y){ double x=y; double s; int n; const sincos_t*p=&T[0];
 if(abstop12(y)<abstop12(0x1.921FB6p-1f)){ s=x*x; if(abstop12(y)<abstop12(0x1p-12f)) return y; return poly(x,s,p,0);} 
 else if(abstop12(y)<abstop12(120.0f)){ x=reduce_fast(x,p,&n); s=p->sign[n&3]; if(n&2)p=&T[1]; return poly(x*s,x*x,p,n);} 
 else if(abstop12(y)<abstop12(INFINITY)){ uint32_t xi=asuint(y); int sign=xi>>31; x=reduce_large(xi,&n); s=p->sign[(n+sign)&3]; if((n+sign)&2)p=&T[1]; return poly(x*s,x*x,p,n);} 
 return y-y; }
float my_cosf(float y){ double x=y; double s; int n; const sincos_t*p=&T[0];
 if(abstop12(y)<abstop12(0x1.921FB6p-1f)){ double x2=x*x; if(abstop12(y)<abstop12(0x1p-12f)) return 1.0f; return poly(x,x2,p,1);} 
 else if(abstop12(y)<abstop12(120.0f)){ x=reduce_fast(x,p,&n); s=p->sign[n&3]; if(n&2)p=&T[1]; return poly(x*s,x*x,p,n^1);} 
 else if(abstop12(y)<abstop12(INFINITY)){ uint32_t xi=asuint(y); int sign=xi>>31; x=reduce_large(xi,&n); s=p->sign[(n+sign)&3]; if((n+sign)&2)p=&T[1]; return poly(x*s,x*x,p,n^1);} 
 return y-y; }
int main(int argc,char**argv){ uint64_t lo=strtoull(argv[1],0,0), hi=strtoull(argv[2],0,0); uint64_t bad_s=0,bad_c=0; 
 for(uint64_t u=lo;u<hi;u++){ uint32_t v=(uint32_t)u; float f; memcpy(&f,&v,4); if(f!=f) continue; if(isinf(f)) continue; float a=sinf(f), b=my_sinf(f); if(asuint(a)!=asuint(b)){ if(bad_s<5) printf("sin %a glibc %a mine %a\n",f,a,b); bad_s++;} a=cosf(f); b=my_cosf(f); if(asuint(a)!=asuint(b)){ if(bad_c<5) printf("cos %a glibc %a mine %a\n",f,a,b); bad_c++;} }
 printf("range %llx-%llx bad sin %llu cos %llu\n",(unsigned long long)lo,(unsigned long long)hi,(unsigned long long)bad_s,(unsigned long long)bad_c); return 0;}
