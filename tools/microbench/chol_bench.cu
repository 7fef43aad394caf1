#include <cstdio>
#include <cuda_runtime.h>
#define PK(i,j) ((i)*((i)+1)/2+(j))
template<int NC>
__device__ __noinline__ void chol(double* A, double* dinv, int n, int t) {
    double a[NC];
    const bool row = t <= n;
    const double* mine = A + PK(t,0);
#pragma unroll
    for (int j=0;j<NC;++j) a[j] = (row && j<n && (j<=t || t==n)) ? mine[j] : 0.0;
#pragma unroll
    for (int k=0;k<NC;++k) {
        if (k<n) {
            double d = __shfl_sync(0xffffffffu, a[k], k);
            if (!(d>0.0)) {d=1.0;}
            const double rs = rsqrt(d);
            const double l = a[k]*rs;
            if (row && t>k) A[PK(t,k)] = l;
            if (t==k) dinv[k]=rs;
            __syncwarp();
#pragma unroll
            for (int j=k+1;j<NC;++j) if (j<n) a[j] -= l*A[PK(j,k)];
        }
    }
    __syncwarp();
}
template<int NC>
__global__ void k1(const double* Ag, double* out, int n, long long* cyc, int reps) {
    extern __shared__ double sm[];
    int w = threadIdx.x/32, t = threadIdx.x%32;
    constexpr int SZ=(NC+1)*(NC+2)/2;
    double* A = sm + w*(SZ+NC); double* dinv = A+SZ;
    long long tot=0;
    for (int rep=0; rep<reps; ++rep) {
    for (int i=t;i<SZ;i+=32) A[i]=Ag[i];
    __syncwarp();
    long long t0=clock64();
    chol<NC>(A, dinv, n, t);
    tot += clock64()-t0;
    }
    if (t==0) { cyc[blockIdx.x*(blockDim.x/32)+w]=tot/reps; out[blockIdx.x]=A[PK(n,n-1)]+dinv[3]; }
}
int main(){
    const int NC=30, SZ=(NC+1)*(NC+2)/2;
    double h[SZ];
    for(int i=0;i<=NC;i++) for(int j=0;j<=i;j++) h[PK(i,j)] = (i==j)? 40.0+i : 0.3+0.01*(i+j);
    double *dA,*dout; long long* dc; cudaMalloc(&dA,sizeof(h)); cudaMalloc(&dout,8*4096); cudaMalloc(&dc,8*65536);
    cudaMemcpy(dA,h,sizeof(h),cudaMemcpyHostToDevice);
    for (int wpb : {1,4}) for (int blocks : {148, 148*4}) {
        size_t smem = wpb*(SZ+NC)*8;
        cudaFuncSetAttribute(k1<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k1<NC><<<blocks, 32*wpb, smem>>>(dA,dout,NC,dc,20);
        cudaDeviceSynchronize();
        long long hc[8]; cudaMemcpy(hc,dc,64,cudaMemcpyDeviceToHost);
        printf("warps/block %d blocks %d (=%d warps/SM): cycles per cholesky %lld  err %s\n", wpb, blocks, wpb*blocks/148, hc[0], cudaGetErrorString(cudaGetLastError()));
    }
}
