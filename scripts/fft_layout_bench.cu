// cuFFT Z2Z 3-D throughput for the layouts considered in DESIGN.md: contiguous batch vs band-interleaved.
#include <cufft.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do{auto e=(x); if(e){printf("err %d line %d\n",(int)e,__LINE__); return 1;}}while(0)
int main(int argc, char** argv){
  int n0=90,n1=90,n2=90, nb=512;
  if(argc>3){n0=atoi(argv[1]);n1=atoi(argv[2]);n2=atoi(argv[3]);}
  if(argc>4) nb=atoi(argv[4]);
  long N=(long)n0*n1*n2;
  cufftDoubleComplex* x; CK(cudaMalloc(&x, sizeof(cufftDoubleComplex)*N*nb)); CK(cudaMemset(x,0,sizeof(cufftDoubleComplex)*N*nb));
  void* work=nullptr; size_t wsmax=0;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  long long nn[3]={n0,n1,n2};
  auto run=[&](const char* name, long long istride, long long idist, int batch, int reps_groups)->int{
    cufftHandle h; CK(cufftCreate(&h)); CK(cufftSetAutoAllocation(h,0)); size_t ws=0;
    long long emb[3]={n0,n1,n2};
    auto r=cufftMakePlanMany64(h,3,nn,emb,istride,idist,emb,istride,idist,CUFFT_Z2Z,batch,&ws);
    if(r){printf("%s: plan failed %d\n",name,(int)r); return 0;}
    if(ws>wsmax){ if(work) cudaFree(work); CK(cudaMalloc(&work,ws)); wsmax=ws;}
    CK(cufftSetWorkArea(h,work));
    for(int w=0;w<2;w++) for(int g=0;g<reps_groups;g++) CK(cufftExecZ2Z(h,x+(long)g*(istride==1? (long)batch*N : batch),x+(long)g*(istride==1?(long)batch*N:batch),CUFFT_INVERSE));
    cudaEventRecord(e0);
    int it=5;
    for(int i=0;i<it;i++) for(int g=0;g<reps_groups;g++) CK(cufftExecZ2Z(h,x+(long)g*(istride==1? (long)batch*N : batch),x+(long)g*(istride==1?(long)batch*N:batch),CUFFT_INVERSE));
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=it;
    double boxes=(double)batch*reps_groups;
    printf("%-40s %8.3f ms for %4.0f boxes  %7.2f us/box  alg %.0f GB/s  (work %zu MB)\n",name,ms,boxes,ms*1e3/boxes,32.0*N*boxes/(ms*1e-3)/1e9, ws>>20);
    cufftDestroy(h); return 0;
  };
  printf("grid %dx%dx%d nb=%d\n",n0,n1,n2,nb);
  run("contiguous batch=nb", 1, N, nb, 1);
  run("contiguous batch=64 x groups", 1, N, 64, nb/64);
  run("contiguous batch=1 x nb", 1, N, 1, nb);
  // interleaved: element (g, b) at g*nb + b : stride nb, dist 1, batch nb
  run("interleaved stride=nb batch=nb", nb, 1, nb, 1);
  // interleaved in groups of 32: x[grp][g][32]
  for(int B: {8,16,32,64}){ char nm[64]; sprintf(nm,"interleaved groups of %d",B);
    // group g starts at g*B*N ; stride B dist 1 batch B
    cufftHandle h; cufftCreate(&h); cufftSetAutoAllocation(h,0); size_t ws=0; long long emb[3]={n0,n1,n2};
    auto r=cufftMakePlanMany64(h,3,nn,emb,B,1,emb,B,1,CUFFT_Z2Z,B,&ws);
    if(r){printf("%s plan failed\n",nm); continue;}
    if(ws>wsmax){ if(work) cudaFree(work); cudaMalloc(&work,ws); wsmax=ws;}
    cufftSetWorkArea(h,work);
    int groups=nb/B;
    for(int g=0;g<groups;g++) cufftExecZ2Z(h,x+(long)g*B*N,x+(long)g*B*N,CUFFT_INVERSE);
    cudaEventRecord(e0); int it=3;
    for(int i=0;i<it;i++) for(int g=0;g<groups;g++) cufftExecZ2Z(h,x+(long)g*B*N,x+(long)g*B*N,CUFFT_INVERSE);
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=it;
    printf("%-40s %8.3f ms for %4d boxes  %7.2f us/box  alg %.0f GB/s  (work %zu MB)\n",nm,ms,nb,ms*1e3/nb,32.0*N*nb/(ms*1e-3)/1e9, ws>>20);
    cufftDestroy(h);
  }
  return 0;
}
