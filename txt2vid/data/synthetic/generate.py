"""txt2vid.data.synthetic.generate on the device: `MovingDigits` draws the reference generator's random
decisions on the host and renders clips + caption tokens with t2v_moving_digits / t2v_grammar_tokens."""
from txt2vid_b200.data import MovingDigits  # noqa: F401
