"""Analysis script (test infrastructure, CPU): emulates 16-bit GEMM-operand rounding on the fp32 oracle to size the
parity budget quoted in DESIGN.md "Precision".  python tests/precision/<this file>"""
import sys, math, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from cpt_b200 import config as C
from cpt_b200.synthetic import synth_state_dict, synth_batch, synth_vocab_ids
from oracle import cpt_oracle as O
torch.set_num_threads(8)
cfg=C.oscar_base(); sd=synth_state_dict(cfg,88); B=8
b=synth_batch(cfg,B,70,50,88); vids=synth_vocab_ids(cfg,8,88)
with torch.no_grad():
    ref_seq,ref_pool,_=O.bert_img_model(sd,cfg,b['input_ids'],b['token_type_ids'],b['attention_mask'],img_feats=b['img_feats'])
    ref=O.cpt_mlm_logits(sd,cfg,b['input_ids'],b['token_type_ids'],b['attention_mask'],b['img_feats'],b['mask_pos'],vids)
    rows=ref_seq[torch.arange(B),b['mask_pos']]
    ref_full=O.lm_head(sd,cfg,rows)
print('ref logits absmax',ref.abs().max().item(),'full-vocab absmax',ref_full.abs().max().item(),'seq absmax',ref_seq.abs().max().item())

def run(qa,qw,resid16=False,headq=True,split=None):
    # qa: quantizer for activations feeding GEMMs; qw: for weights
    def lin(x,pre,q=True):
        W=sd[pre+'.weight']; bb=sd[pre+'.bias']
        if not q: return F.linear(x,W,bb)
        if split=='x3':
            xh=qa(x); xl=qa(x-xh); Wh=qw(W); Wl=qw(W-Wh)
            return F.linear(xh,Wh)+F.linear(xl,Wh)+F.linear(xh,Wl)+bb
        if split=='a2':
            xh=qa(x); xl=qa(x-xh); Wh=qw(W)
            return F.linear(xh,Wh)+F.linear(xl,Wh)+bb
        return F.linear(qa(x),qw(W),bb)
    with torch.no_grad():
        h=O.text_embeddings(sd,cfg,b['input_ids'],b['token_type_ids'])
        im=lin(b['img_feats'],'bert.img_embedding')
        im=O._ln(im,sd['bert.LayerNorm.weight'],sd['bert.LayerNorm.bias'],cfg.img_layer_norm_eps)
        h=torch.cat((h,im),1)
        ext=O.extended_attention_mask(b['attention_mask'])
        nH=12;dH=64;S=h.shape[1]
        for i in range(12):
            p='bert.encoder.layer.%d.'%i
            res=qa(h) if resid16 else h
            q=qa(lin(h,p+'attention.self.query')).view(B,S,nH,dH).permute(0,2,1,3)
            k=qa(lin(h,p+'attention.self.key')).view(B,S,nH,dH).permute(0,2,1,3)
            v=qa(lin(h,p+'attention.self.value')).view(B,S,nH,dH).permute(0,2,1,3)
            sc=torch.matmul(q,k.transpose(-1,-2))/8.0+ext
            pr=qa(torch.softmax(sc,-1))
            ctx=torch.matmul(pr,v).permute(0,2,1,3).reshape(B,S,768)
            a=lin(ctx,p+'attention.output.dense')+res
            a=O._ln(a,sd[p+'attention.output.LayerNorm.weight'],sd[p+'attention.output.LayerNorm.bias'],1e-12)
            res=qa(a) if resid16 else a
            it=O._gelu(lin(a,p+'intermediate.dense'))
            o=lin(it,p+'output.dense')+res
            h=O._ln(o,sd[p+'output.LayerNorm.weight'],sd[p+'output.LayerNorm.bias'],1e-12)
        rows=h[torch.arange(B),b['mask_pos']]
        lg=O.lm_head(sd,cfg,rows,vids)
    e_seq=(h-ref_seq).abs().max().item()/ref_seq.abs().max().item()
    e_lg=(lg-ref).abs().max().item()
    return e_seq,e_lg/ref.abs().max().item(),e_lg/ref_full.abs().max().item(), ((lg-ref).abs()/(ref.abs())).max().item()

f16=lambda x:x.half().float(); bf=lambda x:x.bfloat16().float()
def tf32(x):
    i=x.view(torch.int32); i=(i+0x1000)&~0x1fff; return i.view(torch.float32)
for name,kw in [('bf16',dict(qa=bf,qw=bf)),('fp16',dict(qa=f16,qw=f16)),('fp16 resid16',dict(qa=f16,qw=f16,resid16=True)),
                ('tf32',dict(qa=tf32,qw=tf32)),('fp16 a2',dict(qa=f16,qw=f16,split='a2')),('bf16x3',dict(qa=bf,qw=bf,split='x3')),('fp16x3',dict(qa=f16,qw=f16,split='x3'))]:
    print(name,'seq relmax %.2e | logits err/max|K logits| %.2e | err/max|full-vocab row| %.2e | elementwise rel %.2e'%run(**kw))
