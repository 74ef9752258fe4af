import os, sys
import torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
pipe.set_backend(ec.EC_BACKEND_OZAKI_I8)
n, batch = 512, 416
A = torch.empty(n*n*batch, dtype=torch.float64, device=dev); B = torch.empty(n*n, dtype=torch.float64, device=dev); Cm = torch.empty(n*n*batch, dtype=torch.float64, device=dev)
pipe.fill_uniform(1, 1, A.data_ptr(), n*n, batch); pipe.fill_uniform(1, 0, B.data_ptr(), n*n, 1)
for _ in range(2):
    pipe.dgemm_strided_batched(0,0,n,n,n,1.0,A.data_ptr(),n,n*n,B.data_ptr(),n,0,0.0,Cm.data_ptr(),n,n*n,batch)
nn = 4096
A2 = torch.empty(nn*nn, dtype=torch.float64, device=dev); B2 = torch.empty(nn*nn, dtype=torch.float64, device=dev); C2 = torch.empty(nn*nn, dtype=torch.float64, device=dev)
pipe.fill_uniform(1, 1, A2.data_ptr(), nn*nn, 1); pipe.fill_uniform(1, 0, B2.data_ptr(), nn*nn, 1)
for _ in range(2):
    pipe.dgemm(0,0,nn,nn,nn,1.0,A2.data_ptr(),nn,B2.data_ptr(),nn,0.0,C2.data_ptr(),nn)
torch.cuda.synchronize()
print("done")
