import sys, types, torch
sys.path.insert(0, ".")
import bench, jclip_b200 as jb
args = bench.parse() if False else types.SimpleNamespace(crops=64, images_per_gpu=128, img_dtype="u8", chunk_views=0)
dev = torch.device("cuda", 0)
sd, model, lora, texts, lp_np = bench.build_problem(jb, torch, dev, args)
lp = jb.Channel_LP(); lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = [torch.from_numpy(a) for a in lp_np]
bank = jb.TextBank(texts[0].to(dev), texts[1].to(dev), texts[2].to(dev), dev)
hp = jb.HotPath(model, bank, lp, rank_by="cs5", k=5)
ctx = jb.get_context(dev)
images = (jb.synth.make_views_torch(1000, 128, 65, dev) * 255).round_().to(torch.uint8)
for _ in range(3): hp.evaluate_base(images, topk_to_host=False)
def run(prof, K=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if prof: ctx.profile_start()
    e0.record()
    for _ in range(K): hp.evaluate_base(images, topk_to_host=False)
    e1.record(); torch.cuda.synchronize()
    if prof: ctx.profile_stop()
    return e0.elapsed_time(e1) / K
for rep in range(3):
    print("rep", rep, "no-prof %.3f ms/step" % run(False), "prof %.3f ms/step" % run(True))
