"""`src.compression.token_importance` of the reference, served by rtkv."""
from rtkv.token_importance import LayerWiseImportanceTracker, PromptGuidedImportanceScorer  # noqa: F401
