#include "../structure_from_motion_b200/csrc/sfm_fastsvd.cuh"
extern "C" int null4(const double* g, double* x) { double a[16], o[4]; for (int i=0;i<16;++i) a[i]=g[i]; bool ok = sfm::null_vector4_fast(a, o); for(int i=0;i<4;++i) x[i]=o[i]; return ok; }
extern "C" int frames3(const double* A, double* U, double* V, double* sv) { double a[9],u[9],v[9],s[3]; for(int i=0;i<9;++i)a[i]=A[i]; bool ok=sfm::svd3_rank2_frames(a,u,v,s); for(int i=0;i<9;++i){U[i]=u[i];V[i]=v[i];} for(int i=0;i<3;++i) sv[i]=s[i]; return ok; }
