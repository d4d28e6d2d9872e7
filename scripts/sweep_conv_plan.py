import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
layer = {'conv42': "('conv42', 50, 10, 1024, 1024, True, True)", 'conv22': "('conv22', 200, 40, 256, 256, True, False)",
         'conv32': "('conv32', 100, 20, 512, 512, True, False)", 'conv21': "('conv21', 200, 40, 128, 256, False, False)",
         'conv12': "('conv12', 400, 80, 128, 128, True, False)", 'conv41': "('conv41', 50, 10, 512, 1024, False, False)",
         'conv31': "('conv31', 100, 20, 256, 512, False, False)"}
src = open(os.path.join(here, 'bench_conv_layers.py')).read()
plans = {'conv42': ['10,10,2', '2,50,2', '10,18,1'], 'conv22': ['20,12,1', '40,6,1', '2,40,3', '8,20,1', '10,12,2', '20,6,2'],
         'conv32': ['4,20,3', '20,12,1', '20,10,1', '10,10,2'], 'conv21': ['10,25,1', '40,6,1', '20,12,1', '8,32,1'],
         'conv12': ['16,16,1', '80,2,1', '40,6,1', '20,12,1'], 'conv41': ['10,25,1', '10,10,2'], 'conv31': ['10,25,1', '20,12,1', '20,10,1']}
for name in sys.argv[1:] or plans:
    code = src.replace(src[src.index('layers = ['):src.index('force_nopool')], 'layers = [%s]\n' % layer[name]).replace('os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', 'os.getcwd()')
    for pl in plans[name]:
        env = dict(os.environ, DASV_CONV_PLAN=pl)
        r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
        print(name, pl, r.stdout.strip() or r.stderr.strip()[-200:], flush=True)
