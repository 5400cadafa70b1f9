#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t su32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p);}
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int z, int bw, int bh, unsigned* out, int* status)
{
    extern __shared__ __align__(128) unsigned char raw[];
    unsigned char* sm = raw + ((128u - (su32(raw) & 127u)) & 127u);
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = su32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"((uint32_t)(bw*bh*4)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            :: "r"(su32(sm)), "l"(&tmap), "r"(x0), "r"(y0), "r"(z), "r"(b) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u<<22); spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
    if (!done) { if (threadIdx.x==0) *status = 1; return; }
    const unsigned* t = (const unsigned*)sm;
    for (int i = threadIdx.x; i < bw*bh; i += blockDim.x) out[i] = t[i];
}
int main(int argc, char** argv){
    const int W=641,H=481,B=2, SP=644; const int bw=100,bh=49;
    std::vector<unsigned> h((size_t)B*H*SP);
    for (size_t i=0;i<h.size();i++) h[i]=(unsigned)i;
    unsigned *d,*o; int* st; cudaMalloc(&d,h.size()*4); cudaMemcpy(d,h.data(),h.size()*4,cudaMemcpyHostToDevice);
    cudaMalloc(&o,bw*bh*4); cudaMalloc(&st,4); cudaMemset(st,0,4);
    void* fn=nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t er = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fn,cudaEnableDefault,&q);
    printf("entry %d q %d fn %p\n",(int)er,(int)q,fn);
    typedef CUresult (*Enc)(CUtensorMap*,CUtensorMapDataType,cuuint32_t,void*,const cuuint64_t*,const cuuint64_t*,const cuuint32_t*,const cuuint32_t*,CUtensorMapInterleave,CUtensorMapSwizzle,CUtensorMapL2promotion,CUtensorMapFloatOOBfill);
    CUtensorMap tm;
    cuuint64_t dims[3]={(cuuint64_t)W,(cuuint64_t)H,(cuuint64_t)B}; cuuint64_t str[2]={(cuuint64_t)SP*4,(cuuint64_t)H*SP*4};
    cuuint32_t box[3]={(cuuint32_t)bw,(cuuint32_t)bh,1}; cuuint32_t es[3]={1,1,1};
    CUresult r=((Enc)fn)(&tm,CU_TENSOR_MAP_DATA_TYPE_UINT32,3,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n",(int)r);
    int xs[] = {9, -16, 8, 600, 8}; int ys[] = {4, 4, -17, 4, 460}; 
    int which = argc > 1 ? atoi(argv[1]) : 0;
    {
        int x0 = xs[which], y0 = ys[which], z = 0;
        k<<<1,256,bw*bh*4+128>>>(tm,x0,y0,z,bw,bh,o,st);
        cudaError_t e2=cudaDeviceSynchronize(); int s=0; cudaMemcpy(&s,st,4,cudaMemcpyDeviceToHost);
        std::vector<unsigned> ho(bw*bh); cudaMemcpy(ho.data(),o,bw*bh*4,cudaMemcpyDeviceToHost);
        int bad=0; for(int yy=0;yy<bh;yy++)for(int xx=0;xx<bw;xx++){int gx=x0+xx,gy=y0+yy; unsigned want=(gx<0||gy<0||gx>=W||gy>=H)?0:h[((size_t)z*H+gy)*SP+gx]; if(ho[yy*bw+xx]!=want)bad++;}
        printf("x0=%d y0=%d: err=%s status=%d mismatches=%d\n",x0,y0,cudaGetErrorString(e2),s,bad);
    }
    return 0;
}
