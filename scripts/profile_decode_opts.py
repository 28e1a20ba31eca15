"""profile_decode.py with decode options preset: python scripts/profile_decode_opts.py key=value[,key=value...] [B] [T] [S]"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transformer_tacotron2_b200 as T  # noqa: E402

opts = [kv.split("=") for kv in sys.argv[1].split(",") if kv]
orig = T.TransformerTTS.load_state_dict


def patched(self, *a, **k):
    r = orig(self, *a, **k)
    for key, val in opts:
        self.set_option(key, int(val))
    return r


T.TransformerTTS.load_state_dict = patched
sys.argv = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_decode.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
