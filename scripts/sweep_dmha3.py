import os, subprocess, sys
code = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sweep_dmha.py')).read().split("code = r'''")[1].split("'''")[0]
for env_add in ({}, {'DASV_DMHA_FB': '1'}, {'DASV_DMHA_FB': '1', 'DASV_DMHA_STAGES': '3'}, {'DASV_DMHA_FB': '1', 'DASV_DMHA_FPS': '16'}, {'DASV_DMHA_FPS': '16', 'DASV_DMHA_STAGES': '3'}):
    env = dict(os.environ); env.update(env_add)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(env_add, (r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
