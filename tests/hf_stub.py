"""A minimal transformers model for driving HF `generate()` around the scorer in tests: its "decoder" returns the seeded
synthetic attention log-probs of huggingface_asr_b200.synthetic (a function of the step only), so HF's real beam search
runs the processor on exactly the inputs the shared harness feeds it."""
import torch
from transformers import GenerationMixin, PretrainedConfig, PreTrainedModel
from transformers.modeling_outputs import CausalLMOutputWithPast

from huggingface_asr_b200.generation import JointCTCAttentionGenerationMixin
from huggingface_asr_b200.synthetic import make_attention_scores


class StubConfig(PretrainedConfig):
    model_type = "ctcps_stub_decoder"

    def __init__(self, vocab_size=64, **kw):
        super().__init__(**kw)
        self.vocab_size = vocab_size
        self.num_hidden_layers = 1


class StubDecoder(JointCTCAttentionGenerationMixin, PreTrainedModel, GenerationMixin):
    config_class = StubConfig

    def __init__(self, config, seed=0, scale=0.5, raw_logits=False):
        super().__init__(config)
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.seed, self.scale, self.raw_logits = seed, scale, raw_logits
        self.prefix_len = 0

    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, use_cache=None, **kw):
        # with a cache HF feeds only the new token; the step index is the number of tokens decoded so far
        n = self.prefix_len + input_ids.shape[1] - 1 if past_key_values is None else self.prefix_len
        rows = input_ids.shape[0]
        lp = make_attention_scores(rows, self.config.vocab_size, n, seed=self.seed, scale=self.scale).to(input_ids.device)
        if self.raw_logits:
            lp = lp * 3.0 + 1.5  # greedy search hands raw logits to the processors
        if past_key_values is not None:
            self.prefix_len += 1
        return CausalLMOutputWithPast(logits=lp.unsqueeze(1).expand(-1, input_ids.shape[1], -1).contiguous(),
                                      past_key_values=past_key_values)

    def generate(self, *a, **k):
        self.prefix_len = 0
        return super().generate(*a, **k)
