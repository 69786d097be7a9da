for g in 32 64 128 256 512 1024; do
JAERO_CFE_GROUP=$g timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('group',$g,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],1),'seg_share',round(r['share_of_step'],3),'cfe_share',round(r['cfe_share_of_step'],3),'cfe_ms_per_step',round(r['cfe_share_of_step']*d['ms_per_step'],1))"
done
