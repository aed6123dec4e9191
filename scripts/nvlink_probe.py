"""probe which NVML interface exposes NVLink byte counters on this box"""
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name in ("NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX"):
    fid = getattr(pynvml, name)
    for scope in (None, 0xFFFFFFFF, 0, 1):
        try:
            arg = [fid] if scope is None else [(fid, scope)]
            v = pynvml.nvmlDeviceGetFieldValues(h, arg)[0]
            print(name, scope, "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal, "ui", v.value.uiVal)
        except Exception as e:  # noqa: BLE001
            print(name, scope, "EXC", repr(e))
for link in range(3):
    try:
        print("link", link, "state", pynvml.nvmlDeviceGetNvLinkState(h, link))
    except Exception as e:  # noqa: BLE001
        print("link", link, "EXC", repr(e))
import subprocess
print(subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], capture_output=True, text=True).stdout[:1500])
print(subprocess.run(["nvidia-smi", "nvlink", "-s", "-i", "0"], capture_output=True, text=True).stdout[:600])
