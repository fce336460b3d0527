"""Oracle restatement of torchtune.models (0.4.0). Test infrastructure only."""
from . import llama3_2  # noqa: F401
