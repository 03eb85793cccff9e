from ..random import Key as PRNGKeyArray  # noqa: F401  (annotation target in ofdft_normflows/utils.py:76,110)
