"""Tree-sharded kNN behind the C ABI (mptg_comm_init / mptg_knn_insert_ids / mptg_knn_query_sharded): every rank's slice
of the wave equals the oracle's answer over ALL points -- indices, distances, counts, ties by global index.
world 1 runs on any GPU box; world 2 / 4 need that many GPUs (gpurun --gpus N) and are skipped otherwise."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_knn_equals_single_structure(tmp_path, world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "sharded_worker.py"), str(r), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, p in enumerate(procs):
        assert p.returncode == 0, outs[r][-3000:]
        verdict = (tmp_path / f"rank{r}.txt").read_text()
        print(f"rank {r}: {verdict}")
        assert verdict.startswith("1 "), verdict
