"""GPU: the example scripts (counterparts of the reference's examples/) run end to end at toy sizes."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_optimize_volume_fits_and_writes_a_loadable_asset(tmp_path):
    """examples/optimize_volume.py (reference :124-262): tomography fit of a 6^3 grid to 3 views, pruning, asset export;
    examples/render_asset.py renders the exported asset."""
    out = str(tmp_path / "vol")
    text = _run(["examples/optimize_volume.py", "--output", out, "--cam_count", "3", "--cam_res", "64", "--ref_spp", "4",
                 "--iterations", "12", "--volprim_count", "6", "--sigmat_lr", "0.01"])
    psnr = [float(l.split("psnr=")[1].split()[0]) for l in text.splitlines() if "psnr=" in l and l.startswith("-- step")]
    assert len(psnr) == 12 and psnr[-1] > psnr[0], psnr
    assert os.path.exists(os.path.join(out, "optimized_asset", "__init__.py"))
    text = _run(["examples/render_asset.py", "--asset", os.path.join(out, "optimized_asset"), "--output", str(tmp_path / "r"), "--spp", "2"])
    names = [l.split(":")[0] for l in text.splitlines() if "(64, 64, 3)" in l]
    assert len(names) == 3 and all(os.path.exists(str(tmp_path / "r" / f"{n}.npy")) for n in names), text
    assert np.load(str(tmp_path / "r" / f"{names[0]}.npy")).shape == (64, 64, 3)


@pytest.mark.parametrize("fused", [False, True], ids=["autograd", "fused_step"])
def test_refine_3dg_dataset_reduces_the_loss(fused):
    """examples/refine_3dg_dataset.py (reference :170-189) through loss.backward() and through training.RefineStep."""
    args = ["examples/refine_3dg_dataset.py", "--primitives", "20000", "--cam_count", "3", "--width", "96", "--height", "64",
            "--iterations", "5"] + (["--fused_step"] if fused else [])
    text = _run(args)
    loss = [float(l.split("loss=")[1].split()[0]) for l in text.splitlines() if "loss=" in l]
    assert len(loss) == 5 and loss[-1] < loss[0], loss


def test_render_3dg_asset_runs(tmp_path):
    text = _run(["examples/render_3dg_asset.py", "--primitives", "20000", "--cam_scale", "0.1", "--output", str(tmp_path / "o")])
    assert text
