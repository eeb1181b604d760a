"""hws -- hardware sampler, multi-GPU / B200 successor of /root/reference/src/tcn/hws."""
from .sampler import FakeNVML, NVMLProvider, Sampler  # noqa: F401
