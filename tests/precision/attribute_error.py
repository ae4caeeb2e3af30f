"""Analysis script (test infrastructure, CPU): emulates 16-bit GEMM-operand rounding on the fp32 oracle to size the
parity budget quoted in DESIGN.md "Precision".  python tests/precision/<this file>"""
import sys, math, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from cpt_b200 import config as C
from cpt_b200.synthetic import synth_state_dict, synth_batch, synth_vocab_ids
from oracle import cpt_oracle as O
torch.set_num_threads(8)
cfg=C.oscar_base(); sd=synth_state_dict(cfg,88); B=4
T,R=165,45
b=synth_batch(cfg,B,T,R,88); vids=synth_vocab_ids(cfg,8,88)
with torch.no_grad():
    ref_seq,ref_pool,_=O.bert_img_model(sd,cfg,b['input_ids'],b['token_type_ids'],b['attention_mask'],img_feats=b['img_feats'])
    rows=ref_seq[torch.arange(B),b['mask_pos']]
    ref_full=O.lm_head(sd,cfg,rows)
f16=lambda x:x.half().float()
idt=lambda x:x
def run(q):
    # q: dict of quantizers for: w, x_qkv, qkv, p, ctx, x_up, inter, img
    g=lambda k:q.get(k,idt)
    def lin(x,pre,qa):
        return F.linear(qa(x),g('w')(sd[pre+'.weight']),sd[pre+'.bias'])
    with torch.no_grad():
        h=O.text_embeddings(sd,cfg,b['input_ids'],b['token_type_ids'])
        im=lin(b['img_feats'],'bert.img_embedding',g('img'))
        im=O._ln(im,sd['bert.LayerNorm.weight'],sd['bert.LayerNorm.bias'],cfg.img_layer_norm_eps)
        h=torch.cat((h,im),1)
        ext=O.extended_attention_mask(b['attention_mask'])
        nH=12;dH=64;S=h.shape[1]
        for i in range(12):
            p='bert.encoder.layer.%d.'%i
            qq=g('qkv')(lin(h,p+'attention.self.query',g('x_qkv'))).view(B,S,nH,dH).permute(0,2,1,3)
            k=g('qkv')(lin(h,p+'attention.self.key',g('x_qkv'))).view(B,S,nH,dH).permute(0,2,1,3)
            v=g('qkv')(lin(h,p+'attention.self.value',g('x_qkv'))).view(B,S,nH,dH).permute(0,2,1,3)
            sc=torch.matmul(qq,k.transpose(-1,-2))/8.0+ext
            pr=g('p')(torch.softmax(sc,-1))
            ctx=torch.matmul(pr,v).permute(0,2,1,3).reshape(B,S,768)
            a=lin(ctx,p+'attention.output.dense',g('ctx'))+h
            a=O._ln(a,sd[p+'attention.output.LayerNorm.weight'],sd[p+'attention.output.LayerNorm.bias'],1e-12)
            it=O._gelu(lin(a,p+'intermediate.dense',g('x_up')))
            o=lin(it,p+'output.dense',g('inter'))+a
            h=O._ln(o,sd[p+'output.LayerNorm.weight'],sd[p+'output.LayerNorm.bias'],1e-12)
        pooled=torch.tanh(F.linear(h[:,0],sd['bert.pooler.dense.weight'],sd['bert.pooler.dense.bias']))
        rows=h[torch.arange(B),b['mask_pos']]
        lg=O.lm_head(sd,cfg,rows)
    return ((h-ref_seq).abs().max()/ref_seq.abs().max()).item(), ((lg-ref_full).abs().max(1).values/ref_full.abs().max(1).values).max().item(), (pooled-ref_pool).abs().max().item()
allq={k:f16 for k in ('w','x_qkv','qkv','p','ctx','x_up','inter','img')}
print('all      seq %.2e logits %.2e pooled %.2e'%run(allq))
for k in allq:
    print('%-8s seq %.2e logits %.2e pooled %.2e'%((k,)+run({k:f16})))
for drop in (('w',),('x_qkv','x_up'),('w','x_qkv','x_up'),('inter','ctx'),('w','inter')):
    q=dict(allq)
    for d in drop: q.pop(d)
    print('all but %-18s seq %.2e logits %.2e pooled %.2e'%((','.join(drop),)+run(q)))
