import sys, torch
sys.path.insert(0, ".")
import eigencuda_b200 as ec
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
pipe = ec.CudaPipeline(0, stream=stream.cuda_stream)
nn = 4096
A = torch.empty(nn*nn, dtype=torch.float64, device=dev); B = torch.empty(nn*nn, dtype=torch.float64, device=dev); C = torch.empty(nn*nn, dtype=torch.float64, device=dev)
pipe.fill_uniform(1, 1, A.data_ptr(), nn*nn, 1); pipe.fill_uniform(1, 0, B.data_ptr(), nn*nn, 1)
for backend in (ec.EC_BACKEND_FP64_DMMA, ec.EC_BACKEND_OZAKI_I8):
    pipe.set_backend(backend)
    for _ in range(2):
        pipe.dgemm(0,0,nn,nn,nn,1.0,A.data_ptr(),nn,B.data_ptr(),nn,0.0,C.data_ptr(),nn)
torch.cuda.synchronize(); print("done")
