"""Core API: projection images, lenses, rotation and the (lazy) coordinate map.

Vocabulary (same as the reference's photonbend/core/__init__.py):

* **image** -- uint8 array (height, width, channels).
* **coordinate map** -- for every pixel of an image the ray it looks along, as
  (latitude, longitude, invalid): latitude is the angle from the optical axis (+y, the image
  centre), longitude the angle around it in (-pi, pi], invalid != 0 marks pixels outside the
  image's field of view.  Here it is a lazy ``CoordinateMap`` that converts to the float64
  (H, W, 3) array on demand.
* **protocol** -- ``get_coordinate_map()`` on the destination image, optional
  ``Rotation.rotate_coordinate_map()``, ``process_coordinate_map()`` on the source image.
"""

from photonbend_b200.core.coordinate_map import CoordinateMap
from photonbend_b200.core.lens import (
    Lens,
    equidistant,
    equisolid,
    orthographic,
    rectilinear,
    stereographic,
    thoby,
)
from photonbend_b200.core.projection import (
    CameraImage,
    DoubleCameraImage,
    PanoramaImage,
    ProjectionImage,
    map_projection,
)
from photonbend_b200.core.rotation import Rotation

__all__ = [
    "CoordinateMap", "Lens", "equidistant", "equisolid", "orthographic", "rectilinear",
    "stereographic", "thoby", "CameraImage", "DoubleCameraImage", "PanoramaImage",
    "ProjectionImage", "map_projection", "Rotation",
]
