"""IO: FileSystem back-ends and checkpoint formats."""
from .filesystem import FileSystem, FileSystemBuilder, LocalFileSystem  # noqa: F401
