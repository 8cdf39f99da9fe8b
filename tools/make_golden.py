"""Documents where tests/golden/reference_known_answers.json comes from and re-checks it.

The reference (google/vizier) cannot be imported in the build container: JAX/TFP/flax/equinox are
absent and `vizier.pyvizier` needs generated protobuf modules, so golden
vectors cannot be *generated* by running it.  The file holds the known-answer vectors the
reference's own tests pin for this path, transcribed by hand.  When /root/reference is present this
script greps each cited test file for the literal numbers, so a transcription slip is caught.
"""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_known_answers.json')))
if not os.path.isdir(REF):
  print('reference checkout not present; nothing to verify')
  sys.exit(0)
files = {
    'default_warper_case1': 'vizier/_src/algorithms/designers/gp/output_warpers_test.py',
    'half_rank_case1': 'vizier/_src/algorithms/designers/gp/output_warpers_test.py',
    'half_rank_case2': 'vizier/_src/algorithms/designers/gp/output_warpers_test.py',
    'half_rank_case3': 'vizier/_src/algorithms/designers/gp/output_warpers_test.py',
}
bad = 0
for key, rel in files.items():
  text = open(os.path.join(REF, rel)).read()
  for v in g[key]['expected']:
    if isinstance(v, float) and v != int(v):
      if repr(v) not in text and str(v) not in text:
        print(f'{key}: {v!r} not found in {rel}')
        bad += 1
text = open(os.path.join(REF, 'vizier/_src/algorithms/designers/gp/acquisitions_test.py')).read()
for r in ('0.224', '0.26', '0.44'):
  if r not in text:
    print('radius', r, 'not found'); bad += 1
print('transcription check:', 'OK' if bad == 0 else f'{bad} problems')
sys.exit(1 if bad else 0)
