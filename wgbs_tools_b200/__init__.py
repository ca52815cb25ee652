"""wgbs_tools_b200 -- the per-CpG hot path of wgbstools (bam2pat pileup, pat2beta, homog, segment) as
hand-written sm_100a CUDA behind a C ABI (include/wgbs_b200.h).  No CPU fallback."""
__version__ = "0.1.0"
