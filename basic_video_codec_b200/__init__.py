"""basic_video_codec_b200 -- B200-native encoder hot path of dheri/basic_video_codec.

The compute path is the CUDA library libbvc_b200.so (sm_100a only, C ABI in include/bvc.h).  This
package is the host-side mirror of the reference's interface for that path: EncoderConfig /
InputParameters, PFrame / IFrame with encode_mc_q_dct(), encode_video(), plus encode_clip() for
GOP-batched throughput.  There is no CPU fallback: importing works anywhere, but every compute call
raises if the CUDA library or a B200-class GPU is missing.
"""
from .encoder.params import EncoderConfig  # noqa: F401
from .input_parameters import InputParameters  # noqa: F401
from ._lib import Context, BvcError, library_path, load_library  # noqa: F401
from .clip import encode_clip  # noqa: F401
from .decoder import decode_video, decode_video_framewise  # noqa: F401

__all__ = ["EncoderConfig", "InputParameters", "Context", "BvcError", "encode_clip", "decode_video", "decode_video_framewise", "library_path", "load_library"]
