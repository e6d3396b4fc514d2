# Drop-in `guidance` package: only sd_utils is provided here; any other guidance module (if_utils, zero123_utils, ...)
# found in another `guidance/` directory on sys.path (the reference tree) stays importable.
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
