"""Diagnostic: engine vs golden fixtures (reference outputs), per key, plus stage errors vs the oracle."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_util as pu  # noqa: E402


def main():
    names = sys.argv[1:] or list(pu.CASES)
    for name in names:
        case, sd, bm, x, K, idx = pu.build_inputs(name)
        gold = pu.load_golden(name)
        t0 = time.time()
        m = pu.build_engine(case, sd, bm)
        m.finalize()
        torch.cuda.synchronize()
        print(f"== {name}: engine built in {time.time() - t0:.1f}s")
        focal = float(K[:, 0, 0].max())
        if idx is not None:
            out = m(x, idx=idx, K=K, is_training=True)
            keys = [k for k in gold if k != "idx"]
            bad = pu.compare(out, gold, keys, focal=focal, verbose=True)
        else:
            persons = m(x, K=K, det_thresh=0.3, nms_kernel_size=3)
            print(f"  detected {len(persons)} persons (reference {gold['scores'].shape[0]})")
            if len(persons) == gold["scores"].shape[0]:
                got = {k: torch.stack([p[k] for p in persons]) for k in gold}
                bad = pu.compare(got, gold, list(gold), focal=focal, verbose=True)
            else:
                bad = [("count", len(persons), gold["scores"].shape[0])]
        print("  FAIL:" if bad else "  all keys within tolerance", bad if bad else "")
        # stage: backbone features vs the CPU oracle
        from oracle import dinov2_ref
        with torch.no_grad():
            z_ref = dinov2_ref.get_intermediate_layers(x, sd, case["backbone"], "backbone.encoder.")
        z = m.backbone(x).cpu()
        e = (z - z_ref).abs()
        print(f"  backbone z: max|ref|={z_ref.abs().max():.3f} max err={e.max():.3e} mean err={e.mean():.3e}")
        print(f"  launches per forward: {m.last_launch_count()}")


if __name__ == "__main__":
    main()
