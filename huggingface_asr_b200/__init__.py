"""huggingface_asr_b200 -- the sm_100a CTC prefix scorer of BUTSpeechFIT/huggingface_asr's joint
CTC/attention beam search, as a drop-in for src/decoding/ctc_scorer.py.  See DESIGN.md."""
__version__ = "0.1.0"
