import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge
import torch
cm = ge.load_package()
N=256; n=N**3; nnz=cm.poisson3d_nnz(N)
ia=torch.empty(n+1,dtype=torch.int32,device="cuda"); ja=torch.empty(nnz,dtype=torch.int32,device="cuda"); a=torch.empty(nnz,dtype=torch.float64,device="cuda")
cm.gen_poisson3d_device(N,0,n,ia.data_ptr(),ja.data_ptr(),a.data_ptr())
b=torch.ones(n,dtype=torch.float64,device="cuda")
h_ia=torch.empty(n+1,dtype=torch.int32,pin_memory=True); h_ia.copy_(ia)
h_ja=torch.empty(nnz,dtype=torch.int32,pin_memory=True); h_ja.copy_(ja)
h_a=torch.empty(nnz,dtype=torch.float64,pin_memory=True); h_a.copy_(a)
h_b=torch.empty(n,dtype=torch.float64,pin_memory=True); h_b.copy_(b)
h_x=torch.empty(n,dtype=torch.float64,pin_memory=True)
torch.cuda.synchronize()
for k in range(3):
    st=cm.Stats(); dt=C.c_double(0.0); t0=time.time()
    rc=cm.lib.cudamat_bicgstab_host(0,n,nnz,C.cast(h_a.data_ptr(),cm.c_dp),C.cast(h_ia.data_ptr(),cm.c_ip),C.cast(h_ja.data_ptr(),cm.c_ip),None,None,C.cast(h_b.data_ptr(),cm.c_dp),100,1e-10,0,C.cast(h_x.data_ptr(),cm.c_dp),C.byref(dt),C.byref(st))
    print("call",k,"wall",time.time()-t0, "rc",rc, flush=True)
