"""Plays the renderer in tests/test_gpu_parity.py::test_shareable_output_reaches_another_process: a separate
process that received a file descriptor (as Vulkan / OpenGL would through VK_KHR_external_memory_fd /
GL_EXT_memory_object_fd), maps the memory behind it and prints a digest of the animated rows it finds there."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panda3d_b200 import _capi  # noqa: E402

fd, allocated, nbytes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = _capi.Context(0)
ptr = ctx.shareable_import(fd, allocated)
data = ctx.device_read(ptr, nbytes)
print(hashlib.sha256(bytes(data)).hexdigest())
ctx.shareable_free(ptr)
ctx.close()
