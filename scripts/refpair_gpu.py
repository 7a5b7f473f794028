"""Driver-level check on the reference's shipped scans (refdata/, staged by scripts/refdata_4d.py stage): one pair through
PiecewiseICP_pair_call and through PiecewiseICP_4D_call (reference-epoch mode on a two-file folder), with the reference's
segmentation as the plug-in, against the recorded <e>_Direct2Ref_TransMatrix.txt.  python scripts/refpair_gpu.py 8 12 19"""
import os, sys, subprocess, shutil, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np

def read_T(path):
    l = open(path).read().splitlines()
    return np.array([[float(v) for v in l[1 + r].split()] for r in range(4)])

def child(kind, cfg, out, e):
    from pwicp_b200 import host
    if kind == "pair":
        assert host.pair_call(cfg, out)
    else:
        assert host.call_4d(cfg, 0, 2, 0, 0.75)

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])); sys.exit(0)
    import pwicp_b200 as P
    from pwicp_b200 import synth
    scans = os.path.join(ROOT, "refdata", "scans")
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_supervoxel.so")
    for e in [int(v) for v in sys.argv[1:]]:
        for kind in ("pair", "4d"):
            for order in ("msvc", ""):
                out = tempfile.mkdtemp() + "/"
                cfg = out + "cfg.txt"
                if kind == "pair":
                    synth.write_config(cfg, os.path.join(scans, "Epoch_001.pcd"), os.path.join(scans, "Epoch_%03d.pcd" % e),
                                       res=0.005, sv=0.05, dtinit=0.05, dtmin=0.004)
                else:
                    os.makedirs(out + "scans")
                    for k in (1, e):
                        shutil.copy(os.path.join(scans, "Epoch_%03d.pcd" % k), out + "scans/")
                    synth.write_config(cfg, out + "scans", out, res=0.005, sv=0.05, dtinit=0.05, dtmin=0.004)
                env = dict(os.environ, PWICP_SEGMENTER_PLUGIN=ref_so + ":ref_supervoxel_labels")
                if order: env["PWICP_VOXEL_ORDER"] = order
                r = subprocess.run([sys.executable, __file__, "--child", kind, cfg, out, str(e)], env=env, cwd=out,
                                   capture_output=True, text=True, timeout=900)
                if r.returncode:
                    print(e, kind, order, "FAILED", r.stderr[-500:]); continue
                f = out + ("TransMatrix.txt" if kind == "pair" else "%d_Direct2Ref_TransMatrix.txt" % e)
                T = read_T(f)
                Tr = read_T(os.path.join(ROOT, "refdata", "recorded", "%d_Direct2Ref_TransMatrix.txt" % e))
                da = np.abs(P.matrix2angle(T.astype(np.float32)) - P.matrix2angle(Tr.astype(np.float32))).max()
                dt = np.abs(T[:3, 3] - Tr[:3, 3]).max()
                print(f"epoch {e:2d} {kind:4s} voxel order {order or 'stable':6s}: |d angle| {da:.2e} rad, |d transl| {dt:.2e} m", flush=True)
