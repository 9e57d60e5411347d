#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
static void show(const char*tag){FILE*f=fopen("/proc/self/smaps_rollup","r");char l[256];while(fgets(l,256,f))if(strstr(l,"AnonHuge"))printf("%s %s",tag,l);fclose(f);}
int main(){
  size_t B=(size_t)1<<30;
  for(int thp=0;thp<2;thp++)for(int rep=0;rep<3;rep++){
    char*m=mmap(NULL,B+(2<<20),PROT_READ|PROT_WRITE,MAP_PRIVATE|MAP_ANONYMOUS,-1,0);
    char*b=(char*)(((size_t)m+(2<<20)-1)&~((size_t)(2<<20)-1));
    if(thp){int r=madvise(b,B,MADV_HUGEPAGE);if(r)perror("madvise");}
    double t0=now();
    for(size_t o=0;o<B;o+=4096)b[o]=1;
    double dt=now()-t0;
    printf("thp=%d rep=%d touch 1GB: %.1f ms (%.2f GB/s, %.2f us/4K)\n",thp,rep,dt*1e3,1.0/dt/1.0737*1.0737,dt*1e6/(B/4096));
    show("  ");
    t0=now(); memset(b,2,B); dt=now()-t0; printf("  memset again: %.1f ms (%.1f GB/s)\n",dt*1e3,1.0737/dt);
    munmap(m,B+(2<<20));
  }
  return 0;}
