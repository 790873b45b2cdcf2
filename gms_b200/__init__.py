"""gms-b200: B200-native set-intersection hot path behind GraphMineSuite's API."""
from .capi import (DeviceSet, Graph, GmsbError, Shard, SIM_METRICS, TC_VARIANTS, device_count, generate_rmat, generate_uniform,  # noqa: F401
                   launch_count, lib, set_device, set_devices, set_stream, synchronize)
