"""The algorithm of the training kernels (csrc/lstm_train.cu) and of the GEMMs around them (hss/model/_train.py), restated
step by step in torch on the CPU with the kernels' buffer layouts, against ``nn.LSTM`` under autograd -- what the reference's
training step differentiates (main.py:67-82).  Pins the derivation and the indexing conventions (reverse direction, h_prev
shift, in-place gates -> dG) without a GPU; the kernels themselves are checked in tests/test_train_gpu.py."""
import torch

from oracle import lstm_oracle as lo


def test_bptt_restatement_matches_autograd():
    torch.manual_seed(0)
    B,T,F,H = 3,11,5,8
    params,h0,c0 = lo.reference_params(1,F,B,H)
    x = torch.randn(B,T,F); 
    lstm = torch.nn.LSTM(F,H,bidirectional=True,batch_first=True)
    lstm.load_state_dict({k.split(".",1)[1]:v for k,v in params.items() if k.startswith("lstm_1.")})
    xr = x.clone().requires_grad_(True)
    h0r, c0r = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    out,(hn,cn) = lstm(xr,(h0r,c0r))
    d_out = torch.randn_like(out); d_hn = torch.randn_like(hn); d_cn = torch.randn_like(cn)
    (out*d_out).sum().add((hn*d_hn).sum()).add((cn*d_cn).sum()).backward()

    W = [(lstm.weight_ih_l0.detach(), lstm.weight_hh_l0.detach(), lstm.bias_ih_l0.detach(), lstm.bias_hh_l0.detach()),
         (lstm.weight_ih_l0_reverse.detach(), lstm.weight_hh_l0_reverse.detach(), lstm.bias_ih_l0_reverse.detach(), lstm.bias_hh_l0_reverse.detach())]
    x2 = x.reshape(B*T,F)
    gates = torch.stack([torch.addmm(W[d][2]+W[d][3], x2, W[d][0].t()) for d in range(2)])
    G=4*H
    # emulate fwd kernel
    o = torch.empty(B,T,2*H); cells = torch.empty(2,B*T,H); hn2=torch.empty(2,B,H); cn2=torch.empty(2,B,H)
    for d in range(2):
        for b in range(B):
            h=h0[d,b].clone(); c=c0[d,b].clone()
            for step in range(T):
                t = T-1-step if d else step
                a = gates[d,b*T+t] + W[d][1] @ h
                i,f,g,oo = torch.sigmoid(a[:H]),torch.sigmoid(a[H:2*H]),torch.tanh(a[2*H:3*H]),torch.sigmoid(a[3*H:])
                c = f*c+i*g; h = oo*torch.tanh(c)
                gates[d,b*T+t] = torch.cat([i,f,g,oo]); cells[d,b*T+t]=c; o[b,t,d*H:(d+1)*H]=h
            hn2[d,b]=h; cn2[d,b]=c
    assert (o - out.detach()).abs().max() < 1e-6 and (hn2 - hn.detach()).abs().max() < 1e-6 and (cn2 - cn.detach()).abs().max() < 1e-6
    # emulate bwd kernel
    dG = gates.clone(); dh0=torch.empty(2,B,H); dc0=torch.empty(2,B,H)
    for d in range(2):
        for b in range(B):
            dh_s = d_hn[d,b].clone(); dc_s = d_cn[d,b].clone()
            for step in range(T-1,-1,-1):
                t = T-1-step if d else step
                tp = t+1 if d else t-1
                row=b*T+t
                ig,fg,gg,og = dG[d,row,:H].clone(),dG[d,row,H:2*H].clone(),dG[d,row,2*H:3*H].clone(),dG[d,row,3*H:].clone()
                c = cells[d,row]; c_prev = cells[d,b*T+tp] if step>0 else c0[d,b]
                tc=torch.tanh(c); dh = d_out[b,t,d*H:(d+1)*H]+dh_s
                dc = dc_s + dh*og*(1-tc*tc)
                dao = dh*tc*og*(1-og); dai = dc*gg*ig*(1-ig); dag = dc*ig*(1-gg*gg); daf = dc*c_prev*fg*(1-fg)
                dc_s = dc*fg
                dG[d,row] = torch.cat([dai,daf,dag,dao])
                dh_s = dG[d,row] @ W[d][1]
            dh0[d,b]=dh_s; dc0[d,b]=dc_s
    hp_f = torch.cat([h0[0].unsqueeze(1), o[:, :-1, :H]], dim=1).reshape(B*T,H)
    hp_r = torch.cat([o[:, 1:, H:], h0[1].unsqueeze(1)], dim=1).reshape(B*T,H)
    names = [("weight_ih_l0","weight_hh_l0","bias_ih_l0"),("weight_ih_l0_reverse","weight_hh_l0_reverse","bias_ih_l0_reverse")]
    for d,hp in enumerate((hp_f,hp_r)):
        g=dG[d]
        for got, name in ((g.t() @ x2, names[d][0]), (g.t() @ hp, names[d][1]), (g.sum(0), names[d][2])):
            assert (got - getattr(lstm, name).grad).abs().max() < 1e-5, name
    dx = (dG[0]@W[0][0] + dG[1]@W[1][0]).reshape(B,T,F)
    assert (dx - xr.grad).abs().max() < 1e-5 and (dh0 - h0r.grad).abs().max() < 1e-5 and (dc0 - c0r.grad).abs().max() < 1e-5
