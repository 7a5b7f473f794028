"""BASELINE configs[0] / configs[2] on the REFERENCE'S OWN shipped scans.

The reference ships 20 synthetic epochs (data/data_synthetic/syntheticPC_with_transformations), their ground
truth (defined_transformations.txt), the configuration they were registered with (configuration_files/
configuration_4d.txt: Res 0.005, SV 0.05, DTinit 0.05, DTmin 0.004) and the results its Windows build wrote
(results/4DPCReg/TransParameters*.txt, TransPara_AbsError.txt).  This script runs the same 4D call through
libpwicp_host.so on a B200 and prints, per pair mode, our absolute error against the ground truth next to the
error the reference recorded for itself, and the difference between the two sets of estimated parameters.

    python scripts/refdata_4d.py stage      # here (container): copy the scans into refdata/ (git-ignored, travels with gpurun)
    python scripts/refdata_4d.py run        # on the GPU box: writes gpurun_out/refdata_4d_report.txt
    python scripts/refdata_4d.py run --refseg [--modes 0]
                                            # same, with the reference's own supervoxel segmentation (oracle/_ref/
                                            # libref_supervoxel.so, compiled from the reference's codelibrary) registered
                                            # through the segmenter plug-in: per-epoch 4x4 against the recorded files

The patch generator in front of the hot path is a stand-in (cubic cells instead of Lin's supervoxels), so
agreement with the reference's recorded numbers is expected at the level of its own error against the truth, not bit
level; what the table pins is that the hot path fed with real scans lands where the reference lands.
"""
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
REF = "/root/reference"
STAGE = os.path.join(ROOT, "refdata")
MODES = [(0, "Direct2Ref"), (2, "Fixed"), (-1, "Adaptive")]


def stage():
    os.makedirs(os.path.join(STAGE, "scans"), exist_ok=True)
    src = os.path.join(REF, "data/data_synthetic/syntheticPC_with_transformations")
    for f in sorted(os.listdir(src)):
        shutil.copy(os.path.join(src, f), os.path.join(STAGE, "scans", f))
    shutil.copy(os.path.join(REF, "data/data_synthetic/defined_transformations.txt"), STAGE)
    os.makedirs(os.path.join(STAGE, "recorded"), exist_ok=True)
    for f in os.listdir(os.path.join(REF, "results/4DPCReg")):
        shutil.copy(os.path.join(REF, "results/4DPCReg", f), os.path.join(STAGE, "recorded", f))
    print("staged", len(os.listdir(os.path.join(STAGE, "scans"))), "scans into", STAGE)


def table(path, skip=1):
    return np.array([[float(v) for v in l.split()] for l in open(path).read().splitlines()[skip:] if l.strip()])


def read_T(path):
    l = open(path).read().splitlines()
    return np.array([[float(v) for v in l[1 + r].split()] for r in range(4)]), np.array([[float(v) for v in l[16 + r].split()] for r in range(6)])


def run():
    import ctypes as C
    import pwicp_b200 as P            # noqa: F401  (fails loudly without libpwicp.so / a GPU)
    from pwicp_b200 import host, synth
    refseg = "--refseg" in sys.argv
    modes = MODES
    if "--modes" in sys.argv:
        want = [int(v) for v in sys.argv[sys.argv.index("--modes") + 1].split(",")]
        modes = [m for m in MODES if m[0] in want]
    n_ep = int(sys.argv[sys.argv.index("--epochs") + 1]) if "--epochs" in sys.argv else 20      # epochs 1..n_ep
    out_root = os.path.join(ROOT, "gpurun_out", "refdata_4d_refseg" if refseg else "refdata_4d")
    rep = []
    if refseg:
        ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_supervoxel.so"))
        host.lib().pwicp_host_set_segmenter(C.cast(ref.ref_supervoxel_labels, C.c_void_p))
        rep.append("segmentation: the reference's own supervoxels (oracle/_ref/libref_supervoxel.so) through pwicp_host_set_segmenter")
    else:
        rep.append("segmentation: the library's stand-in (cubic cells)")
    rec_err = table(os.path.join(STAGE, "recorded", "TransPara_AbsError.txt"))
    rec_par = table(os.path.join(STAGE, "recorded", "TransParameters_toRef.txt"))
    rep.append("reference's recorded run (mode unknown from the files; results/4DPCReg): abs error vs ground truth, "
               "max over 19 epochs: rot %.2f mgon, transl %.3f mm; mean: rot %.2f mgon, transl %.3f mm"
               % (rec_err[:, :3].max(), rec_err[:, 3:].max(), rec_err[:, :3].mean(), rec_err[:, 3:].mean()))
    os.environ["PWICP_GROUND_TRUTH"] = os.path.join(STAGE, "defined_transformations.txt")
    for mode, tag in modes:
        out = os.path.join(out_root, tag) + "/"
        os.makedirs(out, exist_ok=True)
        cfg = os.path.join(out, "configuration_4d.txt")
        synth.write_config(cfg, os.path.join(STAGE, "scans"), out, res=0.005, sv=0.05, dtinit=0.05, dtmin=0.004)
        os.chdir(out)
        t0 = time.time()
        ok = host.call_4d(cfg, 0, n_ep, mode, 0.75)
        dt = time.time() - t0
        if not ok:
            rep.append(f"{tag}: call failed")
            continue
        err = table(out + "TransPara_AbsError.txt")
        par = table(out + "TransParameters_toRef.txt")
        d = np.abs(par[:, 1:7] - rec_par[:len(par), 1:7])
        rep.append(f"{tag}: {n_ep - 1} pairs in {dt:.1f} s wall (PCD read + patch generation on the host + device loop)")
        rep.append("  ours vs ground truth : max rot %.2f mgon, max transl %.3f mm; mean rot %.2f mgon, mean transl %.3f mm"
                   % (err[:, :3].max(), err[:, 3:].max(), err[:, :3].mean(), err[:, 3:].mean()))
        rep.append("  ours vs recorded ref : max |d rot| %.2f mgon, max |d transl| %.3f mm (toRef parameters)"
                   % (d[:, :3].max() * 1e3, d[:, 3:].max() * 1e3))
        if refseg:
            rep.append("  per-epoch 4x4 against the recorded <epoch>_%s_TransMatrix.txt: max |d angle| [rad], max |d translation| [m], max rel. d sigma" % tag)
            worst = [0.0, 0.0]
            n_ok = 0
            for e in range(2, n_ep + 1):
                T, V = read_T(out + "%d_%s_TransMatrix.txt" % (e, tag))
                Tr, Vr = read_T(os.path.join(STAGE, "recorded", "%d_%s_TransMatrix.txt" % (e, tag)))
                da = np.abs(P.matrix2angle(T.astype(np.float32)) - P.matrix2angle(Tr.astype(np.float32))).max()
                dtr = np.abs(T[:3, 3] - Tr[:3, 3]).max()
                ds = np.abs(np.sqrt(np.diag(V)) / np.sqrt(np.diag(Vr)) - 1).max()
                n_ok += int(da <= 1e-6 and dtr <= 1e-6)
                rep.append("  %5d  %.2e  %.2e  %.1e" % (e, da, dtr, ds))
            rep.append("  pairs within 1e-6 rad / 1e-6 m of the recorded result: %d of %d" % (n_ok, n_ep - 1))
        rep.append("  epoch  ours[Err_Rx Err_Ry Err_Rz mgon | Err_tx Err_ty Err_tz mm]   recorded[same]")
        for k in range(len(err)):
            rep.append("  %5d  %7.2f %7.2f %7.2f | %6.3f %6.3f %6.3f    %7.2f %7.2f %7.2f | %6.3f %6.3f %6.3f"
                       % ((k + 2,) + tuple(err[k]) + tuple(rec_err[k])))
    text = "\n".join(rep) + "\n"
    with open(os.path.join(ROOT, "gpurun_out", "refdata_4d_refseg_report.txt" if refseg else "refdata_4d_report.txt"), "w") as f:
        f.write(text)
    print(text)


if __name__ == "__main__":
    {"stage": stage, "run": run}[sys.argv[1]]()
