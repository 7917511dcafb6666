"""--sent txt2vid.models.txt.basic.Seq2Seq (scripts/run_tganv2_cond.sh:20)."""
from txt2vid_b200.text import RecurrentModel, Seq2Seq  # noqa: F401
