#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
int main(){
  for (int n=1;n<=24;n++){
    const float d=(float)n, r=1.0f/d;
    long bad=0; uint32_t firstbad=0;
    #pragma omp parallel for reduction(+:bad)
    for (uint32_t b=0;b<0x7f800000u;b++){
      float x; memcpy(&x,&b,4);
      float q=x*r; float e=fmaf(-d,q,x); float q2=fmaf(e,r,q);
      float ref=x/d;
      if (q2!=ref){ bad++; }
    }
    printf("n=%d mismatches %ld\n",n,bad);
  }
}
