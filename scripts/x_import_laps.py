import sys, time, os
sys.path.insert(0,'/root/repo')
os.environ["FB_VERBOSE"]="1"
import numpy as np, bench, femocs_b200 as fb
nodes,hexs,mk=bench.load_x_mesh(2)
ctx=fb.Context(0)
s=fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0))
t=time.perf_counter(); s.import_mesh(nodes,hexs,mk); print("import %.2f s"%(time.perf_counter()-t), flush=True)
s.setup(-bench.E0,0.0); 
t=time.perf_counter(); s.assemble(True); print("assemble %.3f s"%(time.perf_counter()-t), flush=True)
t=time.perf_counter(); it=s.solve(n_cg=3); print("first solve call (3 it; JDS build) %.2f s"%(time.perf_counter()-t), flush=True)
import subprocess; print(subprocess.run("nproc; lscpu | grep 'Model name'", shell=True, capture_output=True, text=True).stdout)
