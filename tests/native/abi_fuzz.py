"""Random argument lists against the host side of every launch entry point (run by tests/test_cabi.py in a subprocess, on
boxes WITHOUT a GPU only: the stand-in addresses are never dereferenced because no launch can happen there).  Every call
must come back with a negative code (contract violation), 0 (empty problem) or a positive cudaError (no device) -- never
crash, divide by zero or overflow on the way."""
import os
import sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b200_ltx import lib
h=lib.load()
random.seed(int(sys.argv[1]))
A=0x10000
def ptr(): return random.choice([None, A, A+4, A+16, A+2])
def dim(): return random.choice([-1,0,1,7,8,64,100,128,129,2048,6144,2**20,2**31-1])
def ld(): return random.choice([0,1,7,8,64,2048,2**31])
cnt={'neg':0,'zero':0,'pos':0}
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 20000):
    name=random.choice(['gemm','fa_fwd','fa_bwd','norm','rowscale','colsum_groups','rf_noise','guid','adamw','lerp','merge','delta'])
    if name=='gemm':
        rc=h.b200_gemm_bf16(ptr(),ld(),random.randint(0,1),ptr(),ld(),random.randint(0,1),ptr(),ld(),ptr(),ld(),random.choice([0,0,8,64,7]),ptr(),ld(),random.randint(0,1),dim(),dim(),dim(),random.randint(-1,4),ptr(),ptr(),ld(),random.choice([0,1,128,-1]),ptr(),ld(),ptr(),ld(),random.choice([0,64,128,256,96]),random.choice([0,1,2,8,-1]),None)
    elif name=='fa_fwd':
        rc=h.b200_fa_fwd(ptr(),ld(),ptr(),ld(),ptr(),ld(),ptr(),ld(),ptr(),ptr(),dim(),random.choice([0,1,4,32,-1]),dim(),dim(),random.choice([64,64,64,128,0]),0.125,None)
    elif name=='fa_bwd':
        rc=h.b200_fa_bwd(ptr(),ld(),ptr(),ld(),ptr(),ld(),ptr(),ld(),ptr(),ptr(),ptr(),ptr(),ld(),ptr(),ld(),ptr(),ld(),dim(),random.choice([0,1,4,32,-1]),dim(),dim(),random.choice([64,64,128]),0.125,ptr(),random.choice([0,1<<20,-1]),None)
    elif name=='norm':
        rc=h.b200_norm_mod_fwd(ptr(),ld(),ptr(),ld(),ptr(),ptr(),ld(),dim(),dim(),random.choice([0,1,4,-1]),1e-6,random.randint(0,1),None)
    elif name=='rowscale':
        rc=h.b200_rowscale(ptr(),ld(),ptr(),ld(),ptr(),ld(),dim(),dim(),random.choice([0,1,4,-1]),None)
    elif name=='colsum_groups':
        rc=h.b200_colsum_groups(ptr(),ld(),ptr(),ld(),ptr(),dim(),dim(),random.choice([0,1,4,-1]),ptr(),0,None)
    elif name=='rf_noise':
        rc=h.b200_rf_noise(ptr(),ptr(),ptr(),ptr(),ptr(),dim(),dim(),None)
    elif name=='guid':
        rc=h.b200_guidance_step(ptr(),ptr(),ptr(),random.choice([0,1,3,-1]),ptr(),random.randint(0,1),ptr(),ptr(),dim(),dim(),dim(),random.randint(0,1),random.randint(0,1),random.randint(0,1),random.randint(0,1),ptr(),random.choice([0,1<<20]),None)
    elif name=='adamw':
        rc=h.b200_adamw_step(ptr(),ptr(),random.choice([-1,0,1,100]),ptr(),ptr(),ptr(),None)
    elif name=='lerp':
        rc=h.b200_lerp_condition(ptr(),ptr(),ptr(),dim(),dim(),dim(),random.choice([0,1,16,384]),0.85,0.5,random.choice([0,-1,16]),dim(),None)
    elif name=='merge':
        rc=h.b200_attn_merge(ptr(),ld(),ptr(),ptr(),ld(),ptr(),ptr(),ld(),dim(),random.choice([0,1,4]),dim(),random.randint(0,1),None)
    else:
        rc=h.b200_attn_delta_zero(ptr(),ld(),ptr(),ld(),ptr(),ptr(),ld(),dim(),random.choice([0,1,4]),dim(),None)
    cnt['neg' if rc<0 else ('zero' if rc==0 else 'pos')]+=1
print(cnt)
