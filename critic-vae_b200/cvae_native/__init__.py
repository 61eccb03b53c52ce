"""Host-side binding of libcvae.so (the C ABI declared in include/cvae.h)."""
from .binding import lib, check, CvaeError, ConvDesc, WgradDesc, PackJob, stream_ptr  # noqa: F401
