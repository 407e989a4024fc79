"""User-facing Recognizer (API shell of danspeech/Recognizer.py: __init__ :39-80, recognize :82-95,
update_model :97-107, update_decoder :109-130, enable_real_time_streaming :499-533,
disable_real_time_streaming :535-558).

The listening API of the reference -- energy VAD over a stream source, background capture and the ``streaming`` /
``real_time_streaming`` generators (Recognizer.py:133-497, :560-818) -- is host logic and lives in
``listening.PhraseListener``, which this class mixes in; only live microphone capture (PyAudio) is outside the
package's environment.
"""
from .DanSpeechRecognizer import DanSpeechRecognizer
from .errors.recognizer_errors import ModelNotInitialized
from .listening import PhraseListener


class Recognizer(PhraseListener):

    def __init__(self, model=None, lm=None, with_gpu=True, **kwargs):
        self._init_listening()
        self.danspeech_recognizer = DanSpeechRecognizer(with_gpu=with_gpu, **kwargs)
        if model:
            self.update_model(model)
        if lm:
            if not model:
                raise ModelNotInitialized("Trying to initialize language model without also choosing a DanSpeech "
                                          "acoustic model.")
            self.update_decoder(lm=lm)
        self.microphone = None

    def recognize(self, audio_data, show_all=False):
        """numpy audio (raw int16 scale, 16 kHz mono) -> transcript (or all beams if ``show_all``)."""
        return self.danspeech_recognizer.transcribe(audio_data, show_all=show_all)

    def recognize_batch(self, audio_list, show_all=False):
        return self.danspeech_recognizer.transcribe_batch(audio_list, show_all=show_all)

    def recognize_batches(self, batches, show_all=False, merge=None):
        """Several batches back to back: up to ``merge`` consecutive batches share one pass of the model (default: the
        engine's ``batches_in_flight``), and the host staging of the next pass overlaps the GPU work of the current."""
        return self.danspeech_recognizer.transcribe_batches(batches, show_all=show_all, merge=merge)

    def update_model(self, model):
        self.danspeech_recognizer.update_model(model)
        print("DanSpeech model updated to: {0}".format(model.model_name))

    def update_decoder(self, lm=None, alpha=None, beta=None, beam_width=None):
        self.danspeech_recognizer.update_decoder(lm=lm, alpha=alpha, beta=beta, beam_width=beam_width)
        print("DanSpeech decoder updated ")

    def enable_real_time_streaming(self, streaming_model, secondary_model=None, string_parts=True):
        self.update_model(streaming_model)
        self.danspeech_recognizer.enable_streaming(secondary_model, string_parts)
        self.stream = True

    def disable_real_time_streaming(self, keep_secondary_model_loaded=False):
        if not self.stream:
            print("No stream is running for the Recognizer")
            return
        print("Stopping microphone stream...")
        self.stream = False
        if self.stream_thread_stopper is not None:     # no capture thread when chunks were pushed by hand
            self.stream_thread_stopper(wait_for_stop=False)
        self.danspeech_recognizer.disable_streaming(keep_secondary_model=keep_secondary_model_loaded)

    def streaming_transcribe(self, chunk, is_last, is_first):
        """One chunk of the real-time path (what Recognizer.real_time_streaming feeds, Recognizer.py:560-715)."""
        return self.danspeech_recognizer.streaming_transcribe(chunk, is_last=is_last, is_first=is_first)
