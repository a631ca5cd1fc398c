"""A built wheel must carry the python packages (incl. the `hugectr` / `hugectr2onnx` module-name shims), the
kernel sources (so a target machine can rebuild for its toolkit) and -- when built after `build_ext` -- the two
shared libraries.  `python tools_dev/check_wheel.py dist/*.whl [--require-libs]`"""
import sys
import zipfile


def main(argv):
    path = [a for a in argv if a.endswith(".whl")][0]
    names = set(zipfile.ZipFile(path).namelist())
    need = ["hugectr_b200/__init__.py", "hugectr_b200/model.py", "hugectr_b200/csrc/gemm_tc2.cu",
            "hugectr_b200/csrc/host/raw_reader.cpp", "hugectr/__init__.py", "hugectr2onnx/__init__.py"]
    if "--require-libs" in argv:
        need += ["hugectr_b200/lib/libhctr_cuda.so", "hugectr_b200/lib/libhctr_host.so"]
    missing = [n for n in need if n not in names]
    print(f"{path}: {len(names)} files" + (f", MISSING {missing}" if missing else ", ok"))
    return 1 if missing else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
