"""The image edge of demo.py (reference demo.py:56-82) around the fused 8-bit forward.

demo.py restores one image per iteration: PIL RGB -> TF.to_tensor (uint8 HWC -> float CHW / 255) -> .cuda() -> model ->
torch.clamp(0, 1) -> permute -> .cpu() -> img_as_ubyte -> BMP.  `restore_u8` keeps the images 8-bit on both sides of the bus
(1 byte per sample over PCIe instead of 4), batches them, and double-buffers pinned-host staging against compute on two side
streams: H2D of batch i+1 and D2H of batch i-1 overlap the forward of batch i.  File decoding / encoding stays with the
caller (PIL / cv2 are not part of this package).
"""
import torch


class U8Pipeline:
    """Batched uint8 restore with pinned staging.  ``run(images)``: images (N, H, W, C) uint8 on the HOST (numpy array or CPU
    tensor) -> (N, H, W, out_chans) uint8 CPU tensor."""

    def __init__(self, model, batch=64, device=None):
        self.model = model
        self.batch = int(batch)
        self.device = torch.device(device if device is not None else next(model.parameters()).device)
        self._bufs = None

    def _buffers(self, h, w, c, oc):
        key = (h, w, c, oc)
        if self._bufs is None or self._bufs[0] != key:
            mk = lambda ch, dev, pin: [torch.empty(self.batch, h, w, ch, dtype=torch.uint8, device=dev, pin_memory=pin) for _ in range(2)]
            self._bufs = (key, mk(c, "cpu", True), mk(c, self.device, False), mk(oc, self.device, False), mk(oc, "cpu", True))
        return self._bufs[1:]

    @torch.no_grad()
    def run(self, images):
        images = torch.as_tensor(images)
        if images.dtype != torch.uint8 or images.dim() != 4:
            raise RuntimeError("U8Pipeline.run: images must be (N, H, W, C) uint8")
        n, h, w, c = images.shape
        net = getattr(self.model, "swin_unet", self.model)
        oc = net.out_chans
        result = torch.empty(n, h, w, oc, dtype=torch.uint8)
        hin, din, dout, hout = self._buffers(h, w, c, oc)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            up, down = torch.cuda.Stream(), torch.cuda.Stream()
            up_done = [torch.cuda.Event() for _ in range(2)]
            fwd_done = [torch.cuda.Event() for _ in range(2)]
            down_done = [torch.cuda.Event() for _ in range(2)]
            starts = list(range(0, n, self.batch))

            def stage(i):
                s, slot = starts[i], i % 2
                cnt = min(self.batch, n - s)
                if i >= 2:
                    fwd_done[slot].synchronize()      # the forward that read this device slot has finished
                hin[slot][:cnt].copy_(images[s:s + cnt])
                with torch.cuda.stream(up):
                    din[slot][:cnt].copy_(hin[slot][:cnt], non_blocking=True)
                    up_done[slot].record(up)

            def drain(i):
                s, slot = starts[i], i % 2
                cnt = min(self.batch, n - s)
                down_done[slot].synchronize()
                result[s:s + cnt].copy_(hout[slot][:cnt])

            if starts:
                stage(0)
            for i, s in enumerate(starts):
                slot = i % 2
                cnt = min(self.batch, n - s)
                if i + 1 < len(starts):
                    stage(i + 1)
                if i >= 2:
                    drain(i - 2)                      # frees hout[slot] / dout[slot] before they are overwritten
                main.wait_event(up_done[slot])
                net.forward_u8(din[slot][:cnt], out=dout[slot][:cnt])
                fwd_done[slot].record(main)
                down.wait_event(fwd_done[slot])
                with torch.cuda.stream(down):
                    hout[slot][:cnt].copy_(dout[slot][:cnt], non_blocking=True)
                    down_done[slot].record(down)
            for i in range(max(0, len(starts) - 2), len(starts)):
                drain(i)
        return result


def restore_u8(model, images, batch=64):
    """One-shot helper: see U8Pipeline."""
    return U8Pipeline(model, batch=batch).run(images)
