"""The committed notebooks are generated from cell scripts: the (fast) embedding-collection one is re-executed here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_embedding_collection_notebook_executes(tmp_path):
    dst = tmp_path / "nb" / "ec.ipynb"
    os.makedirs(dst.parent)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools_dev", "make_notebook.py"),
                        os.path.join(ROOT, "notebooks", "src", "embedding_collection.py"), str(dst)],
                       capture_output=True, text=True, timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0, r.stderr[-2000:]
    nb = json.load(open(dst))
    code = [c for c in nb["cells"] if c["cell_type"] == "code"]
    assert len(code) == 5 and nb["nbformat"] == 4
    text = "".join("".join(o.get("text", [])) for c in code for o in c["outputs"])
    assert "max |diff| after reload: 0.0" in text and "GPU 7: tables" in text
    committed = json.load(open(os.path.join(ROOT, "notebooks", "embedding_collection.ipynb")))
    assert [c["source"] for c in committed["cells"]] == [c["source"] for c in nb["cells"]], \
        "notebooks/embedding_collection.ipynb is stale: rebuild it with tools_dev/make_notebook.py"
