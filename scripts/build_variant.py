"""Build experimental copies of librbf_b200.so with -D overrides into new_bloom_filter_repo_b200/_exp/ (git-ignored *.so, but shipped to
the GPU box): python scripts/build_variant.py name -DX=1 ...   then   RBF_B200_LIB=new_bloom_filter_repo_b200/_exp/lib_name.so python scripts/kbench.py"""
import sys, subprocess, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from new_bloom_filter_repo_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "new_bloom_filter_repo_b200", "_exp")
os.makedirs(root, exist_ok=True)
so = os.path.join(root, "lib_%s.so" % name)
cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-o", so] + [os.path.join(B.CSRC, f) for f in B.SOURCES] + ["-ldl"]
r = subprocess.run(cmd, capture_output=True, text=True)
out = (r.stdout + r.stderr).splitlines()
print(name, "rc", r.returncode)
if r.returncode: print("\n".join(l for l in out if "error" in l))
for i, l in enumerate(out):
    if "k_query4" in l and "Compiling" in l: print("\n".join(out[i + 1:i + 4]))
