"""Run the reference's OWN `run_eval` (scripts/pairwise_register_eval.py:18-461, unmodified, from oracle/_ref)
on a registration model supplied by the caller.  torchio is absent in this image: the script only uses it for
the `tio.DATA` dictionary key, so the mocked module's attribute serves as that key.  Test infrastructure."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import torch

EVAL_METRICS = ["mse", "softdice", "harddice", "harddiceroi", "jdstd", "jdlessthan0"]


def subject(img, lab, name):
    """One entry of the script's paired loader (scripts/register.py:171-209): torchio-style dictionaries."""
    import torchio as tio
    return {"modality": [name], "img": {tio.DATA: img, "affine": torch.eye(4)[None]},
            "seg": {tio.DATA: lab.long()}}


def run_stock_eval(model, img_f, img_m, lab_f, lab_m, out_dir, device, aligns=("rigid", "affine", "tps_1")):
    from oracle import refshim
    refshim.import_reference_scripts()
    from scripts import pairwise_register_eval as pe
    args = SimpleNamespace(early_stop_eval_subjects=None, model_eval_dir=Path(out_dir), skip_if_completed=False,
                           seg_available=True, device=device, num_resolutions_for_itkelastix=None, visualize=False,
                           dim=3, use_amp=False, batch_size=1, save_dir=str(out_dir))
    loader = [(subject(img_f, lab_f, "img_m/IXI_001"), subject(img_m, lab_m, "img_m/IXI_002"))]
    names = [("img_m/IXI_001", "img_m/IXI_002")]
    metrics = pe.run_eval(loader, model, EVAL_METRICS, names, ["rot0"], list(aligns), args)
    save_dir = Path(out_dir) / "eval" / "0_img_m-IXI_001_img_m-IXI_002"
    return metrics, save_dir


def expected_files(aligns=("rigid", "affine", "tps_1")):
    """File set of scripts/pairwise_register_eval.py:368-458 for pair 0, aug rot0."""
    m1, m2 = "img_m-IXI_001", "img_m-IXI_002"
    names = {f"img_f_0-{m1}.npy", f"img_m_0-{m2}-rot0.npy", f"seg_f_0-{m1}.npy", f"seg_m_0-{m2}-rot0.npy",
             f"points_f_0-{m1}.npy", f"points_m_0-{m2}-rot0.npy"}
    for a in aligns:
        pair = f"0-{m1}-{m2}-rot0-{a}"
        names |= {f"metrics-rot0-{a}.json", f"img_a_{pair}.npy", f"grid_{pair}.npy", f"seg_a_{pair}.npy",
                  f"points_a_{pair}.npy"}
    return names
