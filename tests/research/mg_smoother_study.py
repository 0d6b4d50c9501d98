"""CPU study (tests/research, not product code): damped-Jacobi settings of the V-cycle on the bench problem; usage: python tests/research/mg_smoother_study.py <elements per side>.  Results: profiles/research/mg_smoother_study_1024.jsonl"""
import sys, time, json, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import os; HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0,os.path.join(HERE,'..','..')); sys.path.insert(0,HERE)
import mg_prototype as mg
ne=int(sys.argv[1])
t0=time.time()
ops,Ps,b=mg.hierarchy_rediscretized(ne)
A=ops[0]; n=A.shape[0]
print("built",n,len(ops),round(time.time()-t0,1),flush=True)
for nu,om in ((2,0.8),(2,0.67),(2,0.6),(1,0.67),(3,0.67)):
    it=[0]; cnt=[0]*len(ops)
    Am=spla.LinearOperator((n,n),matvec=lambda v:(it.__setitem__(0,it[0]+1),A@v)[1])
    M=spla.LinearOperator((n,n),matvec=lambda v: mg.vcycle(ops,Ps,v,nu=nu,omega=om,counter=cnt))
    x,info=spla.bicgstab(Am,b,rtol=1e-10,atol=0.0,maxiter=150,M=M)
    work=it[0]+sum(c*ops[l].nnz/ops[0].nnz for l,c in enumerate(cnt))
    print(json.dumps({"ne":ne,"nu":nu,"omega":om,"info":info,"outer_spmv":it[0],"fine_spmv_equivalents":round(work,1),
                      "true_rel_res":float(np.linalg.norm(b-A@x)/np.linalg.norm(b))}),flush=True)
