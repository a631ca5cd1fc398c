"""FileSystem abstraction used by every checkpoint read/write.

Parity: HugeCTR/include/io/filesystem.hpp:23-152 (FileSystem ABC: file_size, create_dir, delete_file,
fetch/upload, write, read, copy) with Local / HDFS / S3 / GCS back-ends chosen by path prefix
(FileSystemBuilder::build_unique_by_path).  Remote back-ends use pyarrow.fs when the scheme is
available in the environment; otherwise they raise a clear error (the reference needs the
corresponding SDK at build time: ENABLE_HDFS / ENABLE_S3 / ENABLE_GCS).
"""
from __future__ import annotations

import os
import shutil
from typing import Optional

from ..enums import FileSystemType_t


class FileSystem:
    def get_file_size(self, path: str) -> int: raise NotImplementedError
    def create_dir(self, path: str): raise NotImplementedError
    def delete_file(self, path: str): raise NotImplementedError
    def delete_dir(self, path: str): raise NotImplementedError
    def exists(self, path: str) -> bool: raise NotImplementedError
    def list_dir(self, path: str): raise NotImplementedError          # base names of the entries of a directory
    def write(self, path: str, data: bytes, overwrite: bool = True) -> int: raise NotImplementedError
    def read(self, path: str, offset: int = 0, size: Optional[int] = None) -> bytes: raise NotImplementedError
    def copy(self, src: str, dst: str): raise NotImplementedError
    def fetch(self, remote: str, local: str): self.copy(remote, local)
    def upload(self, local: str, remote: str): self.copy(local, remote)

    # convenience used by the checkpoint writers
    def open(self, path: str, mode: str = "rb"):
        raise NotImplementedError


class LocalFileSystem(FileSystem):
    def get_file_size(self, path): return os.path.getsize(path)
    def create_dir(self, path): os.makedirs(path, exist_ok=True)
    def delete_file(self, path):
        if os.path.exists(path):
            os.remove(path)
    def delete_dir(self, path): shutil.rmtree(path, ignore_errors=True)
    def exists(self, path): return os.path.exists(path)
    def list_dir(self, path): return sorted(os.listdir(path or "."))

    def write(self, path, data, overwrite=True):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(path, "wb" if overwrite else "ab") as f:
            return f.write(data)

    def read(self, path, offset=0, size=None):
        with open(path, "rb") as f:
            f.seek(offset)
            return f.read() if size is None else f.read(size)

    def copy(self, src, dst):
        d = os.path.dirname(dst)
        if d:
            os.makedirs(d, exist_ok=True)
        shutil.copyfile(src, dst)

    def open(self, path, mode="rb"):
        if "w" in mode or "a" in mode:
            d = os.path.dirname(path)
            if d:
                os.makedirs(d, exist_ok=True)
        return open(path, mode)


class ArrowFileSystem(FileSystem):
    """HDFS / S3 / GCS through pyarrow.fs (same call surface as the local one)."""

    def __init__(self, kind: FileSystemType_t, server: str = "", port: int = 0, fs=None):
        import pyarrow.fs as pafs
        self.kind = kind
        if fs is not None:          # an already constructed pyarrow FileSystem (custom endpoint, tests)
            self.fs = fs
            return
        try:
            if kind == FileSystemType_t.HDFS:
                self.fs = pafs.HadoopFileSystem(server or "default", port or 0)
            elif kind == FileSystemType_t.S3:
                self.fs = pafs.S3FileSystem()
            elif kind == FileSystemType_t.GCS:
                self.fs = pafs.GcsFileSystem()
            else:
                raise ValueError(kind)
        except Exception as e:  # SDK / service missing
            raise RuntimeError(f"{kind.name} filesystem is not available in this environment: {e}")

    @staticmethod
    def _strip(path):
        for p in ("hdfs://", "s3://", "gs://", "https://", "mock://"):
            if path.startswith(p):
                return path[len(p):]
        return path

    def get_file_size(self, path): return self.fs.get_file_info(self._strip(path)).size
    def create_dir(self, path): self.fs.create_dir(self._strip(path), recursive=True)
    def delete_file(self, path): self.fs.delete_file(self._strip(path))
    def delete_dir(self, path): self.fs.delete_dir(self._strip(path))
    def exists(self, path):
        import pyarrow.fs as pafs
        return self.fs.get_file_info(self._strip(path)).type != pafs.FileType.NotFound

    def list_dir(self, path):
        import pyarrow.fs as pafs
        return sorted(i.base_name for i in self.fs.get_file_info(pafs.FileSelector(self._strip(path), recursive=False)))

    def _out(self, path):
        """output stream; back-ends with real directories (HDFS, local, in-memory) need the parent first,
        object stores do not -- so the directory is only created when the open fails for its absence"""
        p = self._strip(path)
        try:
            return self.fs.open_output_stream(p)
        except (FileNotFoundError, OSError):
            parent = p.rsplit("/", 1)[0] if "/" in p else ""
            if not parent:
                raise
            self.fs.create_dir(parent, recursive=True)
            return self.fs.open_output_stream(p)

    def write(self, path, data, overwrite=True):
        if not overwrite and self.exists(path):
            data = self.read(path) + data            # (object stores have no append)
        with self._out(path) as f:
            f.write(data)
        return len(data)

    def read(self, path, offset=0, size=None):
        with self.fs.open_input_file(self._strip(path)) as f:
            f.seek(offset)
            return f.read() if size is None else f.read(size)

    def copy(self, src, dst):
        try:
            self.fs.copy_file(self._strip(src), self._strip(dst))
        except (FileNotFoundError, OSError):
            d = self._strip(dst)
            if "/" not in d:
                raise
            self.fs.create_dir(d.rsplit("/", 1)[0], recursive=True)
            self.fs.copy_file(self._strip(src), d)

    def fetch(self, remote, local):
        """remote -> local disk (filesystem.hpp fetch)"""
        d = os.path.dirname(local)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(local, "wb") as f:
            f.write(self.read(remote))

    def upload(self, local, remote):
        with open(local, "rb") as f:
            self.write(remote, f.read())

    def open(self, path, mode="rb"):
        if "r" in mode:
            return self.fs.open_input_file(self._strip(path))
        return self._out(path)


class FileSystemBuilder:
    _registered: dict = {}

    @staticmethod
    def register(scheme: str, arrow_fs) -> None:
        """route paths starting with ``scheme`` (e.g. "mock://", "s3://my-minio/") to a constructed
        ``pyarrow.fs.FileSystem`` -- custom endpoints / credentials, and the in-memory file system of the tests"""
        FileSystemBuilder._registered[scheme] = arrow_fs

    @staticmethod
    def unregister(scheme: str) -> None:
        FileSystemBuilder._registered.pop(scheme, None)

    @staticmethod
    def build_by_path(path: str, params=None) -> FileSystem:
        """FileSystemBuilder::build_unique_by_path: scheme decides the back-end."""
        server = getattr(params, "server", "") if params else ""
        port = getattr(params, "port", 0) if params else 0
        for scheme, fs in FileSystemBuilder._registered.items():
            if path.startswith(scheme):
                return ArrowFileSystem(FileSystemType_t.Other, fs=fs)
        if path.startswith("hdfs://"):
            return ArrowFileSystem(FileSystemType_t.HDFS, server, port)
        if path.startswith("s3://") or ".s3." in path:
            return ArrowFileSystem(FileSystemType_t.S3)
        if path.startswith("gs://") or "storage.googleapis.com" in path:
            return ArrowFileSystem(FileSystemType_t.GCS)
        return LocalFileSystem()

    # names of the reference (include/io/filesystem.hpp:125-152)
    build_unique_by_path = build_by_path

    @staticmethod
    def build_by_type(kind: FileSystemType_t, params=None) -> FileSystem:
        if kind in (FileSystemType_t.Local, FileSystemType_t.Other):
            return LocalFileSystem()
        return ArrowFileSystem(kind, getattr(params, "server", ""), getattr(params, "port", 0))


FileSystemBuilder.build_unique_by_type = FileSystemBuilder.build_by_type


# ---------------------------------------------------------------- array files on any back-end
def is_remote(path: str) -> bool:
    """string test only (no back-end is constructed): a registered scheme, hdfs:// s3:// gs://, or an S3 / GCS URL"""
    return any(path.startswith(sc) for sc in FileSystemBuilder._registered) or \
        path.startswith(("hdfs://", "s3://", "gs://")) or ".s3." in path or "storage.googleapis.com" in path


def write_array(path: str, arr, params=None) -> None:
    """raw little-endian dump of ``arr`` (numpy) -- ``tofile`` locally, one object write remotely"""
    import numpy as np
    arr = np.ascontiguousarray(arr)
    if not is_remote(path):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        arr.tofile(path)
    else:
        FileSystemBuilder.build_by_path(path, params).write(path, arr.tobytes())


def read_array(path: str, dtype, params=None):
    import numpy as np
    if not is_remote(path):
        return np.fromfile(path, dtype=dtype)
    return np.frombuffer(FileSystemBuilder.build_by_path(path, params).read(path), dtype=dtype).copy()


def path_exists(path: str, params=None) -> bool:
    if not is_remote(path):
        return os.path.exists(path)
    return FileSystemBuilder.build_by_path(path, params).exists(path)


def path_join(base: str, *names: str) -> str:
    return "/".join([base.rstrip("/")] + list(names)) if "://" in base else os.path.join(base, *names)
