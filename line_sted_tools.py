"""Drop-in shim: `import line_sted_tools` (the name the reference's figure
scripts use) resolves to the B200 backend when the repo root is first on
sys.path.  See INTEGRATION.md."""
from rescan_line_sted_b200.line_sted_tools import *  # noqa: F401,F403
from rescan_line_sted_b200.line_sted_tools import (  # noqa: F401
    Deconvolver, generate_psfs, get_width, logarithmic_progress, psf_report,
    psf_report_batch, tune_psf)
