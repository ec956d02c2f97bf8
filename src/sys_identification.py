"""Drop-in module path of the reference (`from src.sys_identification import SystemIdentification`,
reference demo/solo_identification.py:6): re-exports the B200-native implementation."""
from system_identification_b200.sys_identification import SystemIdentification  # noqa: F401
