"""photonbend_b200 -- a B200-native (sm_100a) implementation of photonbend's per-pixel remap
path behind photonbend's own Python API.

    from photonbend_b200.core.projection import CameraImage, DoubleCameraImage, PanoramaImage
    from photonbend_b200.core.rotation import Rotation
    from photonbend_b200.core.lens import equidistant
    from photonbend_b200.utils import to_radians

    src = CameraImage(photo, to_radians(360), equidistant(), magnitude=photo.shape[1] / 2 - 0.5)
    dst = PanoramaImage(np.zeros((h, 2 * h, 3), np.uint8))
    cmap = dst.get_coordinate_map()                           # lazy: nothing is computed
    cmap = Rotation(pitch, yaw, roll).rotate_coordinate_map(cmap)   # still lazy
    pano = src.process_coordinate_map(cmap)                   # one fused CUDA kernel

Layers: ``core`` / ``utils`` / ``scripts`` mirror the reference's package; ``engine`` moves
buffers and calls ``libpbremap.so`` (``csrc/``, C ABI in ``include/pb_remap.h``) through ctypes.
There is no CPU fallback: without a CUDA device or the built library every remap call raises.
"""

__version__ = "0.1.0"
