import sys, os, json
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import golden_util as gu
from sickle_b200 import capi, runner
g=json.load(open(os.path.join(ROOT,'tests/golden/golden.json'))); gdir=os.path.join(ROOT,'tests/golden')
modes={"se":capi.MODE_SE,"pei":capi.MODE_PE_INTER,"pe2":capi.MODE_PE_2FILE}
for case in g['cases']:
    if case['threads']>1: continue
    kind,in0,in1=gu.load_inputs(case,gdir); f=gu.parse_flags(case['flags'])
    p=capi.make_params(f['qualtype'],f['q'],f['l'],f['x'],f['n'],mode=modes[kind],has_singles='-s' in case['outputs'])
    try:
        with capi.Context(p,1<<16,1) as ctx:
            r=runner.trim_stream(ctx,in0,in1)
    except runner.DataError as e:
        continue
    except Exception as e:
        print("FAIL", case['id'], repr(e)[:300]); break
print("done")
