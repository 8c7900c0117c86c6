from .api import VESCALE_DEVICE_MESH, VeDeviceMesh  # noqa: F401
