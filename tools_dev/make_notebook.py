"""Build an executed Jupyter notebook (nbformat 4 JSON) from a plain-text cell script, without jupyter:

  python tools_dev/make_notebook.py notebooks/src/quickstart.py notebooks/quickstart.ipynb

Cell script format: lines starting with `# %% [markdown]` open a markdown cell (its lines are comment lines, the
leading "# " is stripped); `# %%` opens a code cell.  Code cells are exec'd in one namespace in order, stdout is
captured into the cell's output (long outputs are trimmed), so the committed notebook shows real results.
"""
import contextlib
import io
import json
import os
import sys
import traceback


def parse(text):
    cells, cur = [], None
    for line in text.splitlines():
        if line.startswith("# %% [markdown]"):
            cur = {"type": "markdown", "src": []}
            cells.append(cur)
        elif line.startswith("# %%"):
            cur = {"type": "code", "src": []}
            cells.append(cur)
        elif cur is not None:
            cur["src"].append(line)
    for c in cells:
        while c["src"] and not c["src"][-1].strip():
            c["src"].pop()
        if c["type"] == "markdown":
            c["src"] = [l[2:] if l.startswith("# ") else l.lstrip("#") for l in c["src"]]
    return [c for c in cells if c["src"]]


def run(cells, cwd):
    ns = {"__name__": "__main__"}
    out_cells, n = [], 0
    old = os.getcwd()
    os.chdir(cwd)
    try:
        for c in cells:
            src = "\n".join(c["src"])
            if c["type"] == "markdown":
                out_cells.append({"cell_type": "markdown", "metadata": {}, "source": src.splitlines(keepends=True)})
                continue
            n += 1
            buf = io.StringIO()
            outputs = []
            with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
                try:
                    exec(compile(src, f"<cell {n}>", "exec"), ns)
                except Exception:
                    outputs.append({"output_type": "error", "ename": "Exception", "evalue": "",
                                    "traceback": traceback.format_exc().splitlines()})
                    raise
                finally:
                    txt = buf.getvalue()
                    lines = txt.splitlines(keepends=True)
                    if len(lines) > 60:
                        lines = lines[:40] + [f"... ({len(lines) - 55} lines trimmed) ...\n"] + lines[-15:]
                    if lines:
                        outputs.insert(0, {"output_type": "stream", "name": "stdout", "text": lines})
            out_cells.append({"cell_type": "code", "execution_count": n, "metadata": {}, "outputs": outputs,
                              "source": src.splitlines(keepends=True)})
    finally:
        os.chdir(old)
    return out_cells


def main(argv):
    src, dst = argv[0], argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    work = os.path.join(os.path.dirname(os.path.abspath(dst)), ".work")
    os.makedirs(work, exist_ok=True)
    cells = run(parse(open(src).read()), work)
    nb = {"cells": cells, "nbformat": 4, "nbformat_minor": 5,
          "metadata": {"kernelspec": {"display_name": "Python 3", "language": "python", "name": "python3"},
                       "language_info": {"name": "python"}}}
    json.dump(nb, open(dst, "w"), indent=1)
    print(f"{dst}: {len(cells)} cells")


if __name__ == "__main__":
    main(sys.argv[1:])
