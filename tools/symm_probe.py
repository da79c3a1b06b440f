import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda",local))
t = symm_mem.empty(1024, dtype=torch.float32, device=torch.device("cuda",local))
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
t.fill_(rank+1)
hdl.barrier()
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
peer=(rank+1)%world
pb = hdl.get_buffer(peer, (1024,), torch.float32)
print(rank, "peer value", float(pb[0].item()), "attrs", [a for a in dir(hdl) if not a.startswith('_')], flush=True)
hdl.barrier()
dist.destroy_process_group()
