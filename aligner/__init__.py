"""`python -m aligner command=evaluate encoder=... data=...` -- the reference's command line, served by the
Lightning/Hydra-free runner in fitclip_b200.runner (the reference's own entry point is aligner/__main__.py:27-93)."""
